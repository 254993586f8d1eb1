"""World-size-2 gloo test of the channel-sharding host logic (unires_b200/parallel.py): the
per-rank arithmetic is the CPU oracle's, the collectives are the ones `_update_admm_sharded`
issues on NCCL.  Result must equal the single-process oracle ADMM iteration."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.nitorch_shim import spatial as S
from oracle import unires_port as P
from tests import _util as U
from unires_b200 import parallel


def test_channel_shard_partition():
    for C in (1, 3, 8, 11):
        for W in (1, 2, 4, 8):
            if W > C:  # a rank without channels is rejected on every rank (no hang in a collective)
                with pytest.raises(ValueError):
                    parallel.channel_shard(C, W, 0)
                continue
            shards = [parallel.channel_shard(C, W, r) for r in range(W)]
            assert sorted(c for s in shards for c in s) == list(range(C))
            assert all(parallel.owner(c, W) == r for r, s in enumerate(shards) for c in s)
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        parallel.channel_shard(3, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _scaled_grad(yc, vx, alpha, z_old_c):
    g = yc.lam * S.im_gradient(yc.dat, vx=vx)
    return g if alpha == 1 else alpha * g + (1 - alpha) * z_old_c


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        _, recipe = U.load_golden('sr3_thick_xyz')
        sc = U.build(recipe, *U.port_namespaces())
        mine = parallel.channel_shard(len(sc.x), world, rank)
        x = [sc.x[c] for c in mine]
        y = [sc.y[c] for c in mine]
        rho, sett = sc.rho, sc.sett
        vx = S.voxel_size(sc.y[0].mat).float()
        z = torch.zeros((len(mine), 3) + tuple(sc.y[0].dim))
        w = torch.zeros_like(z)
        tmp = torch.zeros(sc.y[0].dim)
        # y-update of the local channels only: no collective (the single-rank port does the same)
        zz, ww = P.admm_aux(y)
        P.update_admm(x, y, zz, ww, rho, tmp, torch.zeros(1, 3, dtype=torch.float64), 0,
                      P.Settings(**dict(vars(sett), tolerance=0)))
        # (the call above also ran a *local* prox on zz/ww, which we discard: z, w stay zero)
        row = torch.zeros(3, dtype=torch.float64)
        field = torch.zeros(sc.y[0].dim)

        def data_and_prior(r, f):
            nll = P.compute_nll(x, y, sett)
            r[1] = nll[1]
            f.zero_()
            for yc in y:
                f += torch.sum((yc.lam * S.im_gradient(yc.dat, vx=vx)) ** 2, dim=0)

        def norm2(f):
            f.zero_()
            for k, yc in enumerate(y):
                f += torch.sum((w[k] / rho + _scaled_grad(yc, vx, 1.0, z[k])) ** 2, dim=0)

        jtv = torch.zeros(sc.y[0].dim)

        def apply(f):
            s = f.sqrt()
            fac = (s - 1 / rho).clamp_min(0) / (s + 1e-7)
            jtv.copy_(fac)
            for k, yc in enumerate(y):
                g = _scaled_grad(yc, vx, 1.0, z[k])
                z[k] = fac * (w[k] / rho + g)
                w[k] += rho * (g - z[k])

        # the product's sharded iteration: both fields in ONE all-reduce
        fields = torch.zeros((2,) + tuple(sc.y[0].dim))
        parallel.coupled_objective_and_prox(
            row, fields, data_and_prior, norm2,
            lambda f: torch.sum(torch.sqrt(f), dtype=torch.float64), apply)
        torch.save({'mine': mine, 'y': [yc.dat for yc in y], 'z': z, 'w': w, 'row': row, 'jtv': jtv},
                   os.path.join(out_dir, 'rank%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_sharded_admm_iteration_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    _, recipe = U.load_golden('sr3_thick_xyz')
    sc = U.build(recipe, *U.port_namespaces())
    z, w = P.admm_aux(sc.y)
    obj = torch.zeros(1, 3, dtype=torch.float64)
    _, z, w, jtv, obj, _ = P.update_admm(sc.x, sc.y, z, w, sc.rho, torch.zeros(sc.y[0].dim), obj, 0,
                                         sc.sett)
    for rank in range(world):
        got = torch.load(os.path.join(str(tmp_path), 'rank%d.pt' % rank))
        for k, c in enumerate(got['mine']):
            assert torch.equal(got['y'][k], sc.y[c].dat)
            assert U.rel_l2(got['z'][k], z[c]) < 1e-6
            assert U.rel_l2(got['w'][k], w[c]) < 1e-6
        assert U.rel_l2(got['jtv'], jtv) < 1e-6
        # the float32 energy field is summed over channels in rank order, not channel order
        assert torch.allclose(got['row'], obj[0], rtol=1e-6)
