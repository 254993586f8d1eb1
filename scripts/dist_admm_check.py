"""torchrun --nproc-per-node 2 scripts/dist_admm_check.py
Channel-sharded ADMM iterations on N GPUs (NCCL) against the single-GPU _update_admm."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _project, _update, parallel, struct, synth  # noqa: E402


def main():
    rank, local, world = parallel.env_rank()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    cfg = synth.scaled(synth.CONFIGS['sr3_256'], (64, 72, 128))
    cfg['thick'] = ([(0, 4), (1, 4), (2, 4), (2, 2)] * ((world + 3) // 4))[:max(4, world)]
    full = synth.make_scenario(cfg, _project, struct, device=dev, seed=0)
    C = len(full.x)
    mine = parallel.channel_shard(C, world, rank)
    x = [full.x[c] for c in mine]
    y = [struct._output(dat=full.y[c].dat.clone(), dim=full.y[c].dim, mat=full.y[c].mat,
                        lam=full.y[c].lam) for c in mine]
    z, w = _update._admm_aux(y, full.sett)
    tmp = torch.zeros(full.y[0].dim, device=dev)
    obj = torch.zeros(2, 3, dtype=torch.float64, device=dev)
    for it in range(2):
        _update._update_admm_sharded(x, y, z, w, full.rho, tmp, obj, it, full.sett)
    # reference: all channels on this GPU
    zf, wf = _update._admm_aux(full.y, full.sett)
    tf = torch.zeros(full.y[0].dim, device=dev)
    of = torch.zeros(2, 3, dtype=torch.float64, device=dev)
    for it in range(2):
        _update._update_admm(full.x, full.y, zf, wf, full.rho, tf, of, it, full.sett)
    err = 0.0
    for k, c in enumerate(mine):
        err = max(err, ((y[k].dat - full.y[c].dat).norm() / full.y[c].dat.norm()).item())
        err = max(err, ((z[k] - zf[c]).norm() / zf[c].norm()).item())
        err = max(err, ((w[k] - wf[c]).norm() / wf[c].norm()).item())
    oerr = ((obj - of).abs() / of.abs()).max().item()
    print('rank %d channels %s max rel err %.3e objective rel err %.3e obj %s'
          % (rank, mine, err, oerr, obj[:, 0].tolist()), flush=True)
    assert err < 1e-5 and oerr < 1e-6
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
