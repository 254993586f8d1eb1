"""CG solve and the ADMM iteration on the GPU against the oracle and the golden fixtures:
per-iterate relative L2 <= 1e-4 and identical CG trip counts (north_star)."""
import numpy as np
import pytest
import torch

from oracle import gen_golden
from oracle import unires_port as P
from oracle.nitorch_shim.core import optim as OO
from tests import _util as U

pytestmark = pytest.mark.gpu


def _channel_problem(sc, c, cuda):
    """rhs b and lhs of channel c at z = w = 0 for oracle and product."""
    from unires_b200 import _project
    x, y, sett = U.to_device(sc, cuda)
    vx = torch.ones(3) * float(sc.cfg['vx_y'])
    kw = dict(method=sc.sett.method, do=sc.sett.do_proj)
    b = torch.zeros(sc.y[c].dim)
    for n, obs in enumerate(sc.x[c]):
        b += obs.tau * P.proj('At', obs.dat, sc.x[c], sc.y[c], n=n, **kw)
    lhs_o = lambda v: P.proj('AtA', v, sc.x[c], sc.y[c], rho=sc.rho, vx_y=vx, **kw)
    lhs_g = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj, rho=sc.rho, vx_y=vx)
    return b, lhs_o, lhs_g, y[c].dat


@pytest.mark.parametrize('name', U.GOLDEN_NAMES)
@pytest.mark.parametrize('stop', ['max_gain', 'residual'])
def test_cg_per_iterate_parity_and_trip_count(cuda, name, stop):
    from unires_b200 import optim
    _, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    for c in range(len(sc.x)):
        b, lhs_o, lhs_g, x0 = _channel_problem(sc, c, cuda)
        iterates = {}
        xo = sc.y[c].dat.clone()
        OO.cg(A=lhs_o, b=b, x=xo, max_iter=20, tolerance=1e-3, stop=stop,
              record=lambda it, xi: iterates.__setitem__(it, xi.clone()))
        n_ref, obj_ref = OO.cg.last_n_iter, OO.cg.last_obj
        # full solve with the device-side stop test
        xg = x0.clone()
        optim.cg(A=lhs_g, b=b.to(cuda), x=xg, max_iter=20, tolerance=1e-3, stop=stop)
        info = optim.cg.last
        assert info.n_iter == n_ref, (name, c, stop)
        assert U.rel_l2(xg, xo) < U.REL_TOL
        # the objective trace.  The kernels round D'D differently from the reference
        # (diag*c - sum(neighbours) vs differences of differences); the energy 0.5 x'Ax - b'x and
        # sqrt(r.r) of the recursively updated residual both lose relative accuracy as they
        # shrink towards convergence, hence an absolute floor relative to the first value
        # (measured on B200, worst fixture iso2_1ch, residual rule, iterate 18 of 20: deviation
        # 1.0e-6 |obj[0]| + 1e-4 |obj| with the lean kernel, 1e-7 with the direct kernel whose
        # D'D is formed from differences like the reference's)
        assert np.allclose(info.obj, obj_ref.numpy(), rtol=1e-4,
                           atol=3e-6 * abs(obj_ref[0].item()))
        # per-iterate parity: fixed trip counts, no stop test
        for k in sorted(set([1, 2, 3, n_ref])):
            xk = x0.clone()
            optim.cg(A=lhs_g, b=b.to(cuda), x=xk, max_iter=k, tolerance=0, stop=stop)
            assert optim.cg.last.n_iter == k
            assert U.rel_l2(xk, iterates[k]) < U.REL_TOL, (name, c, k)


def test_cg_generic_callable_path(cuda):
    """cg() over an arbitrary callable uses the same CUDA vector kernels from a host loop."""
    from unires_b200 import optim
    _, recipe = U.load_golden('thickz2_scl')
    sc = U.build(recipe, *U.port_namespaces())
    b, lhs_o, lhs_g, x0 = _channel_problem(sc, 0, cuda)
    xo = sc.y[0].dat.clone()
    OO.cg(A=lhs_o, b=b, x=xo, max_iter=20, tolerance=1e-3, stop='max_gain')
    xg = x0.clone()
    optim.cg(A=lambda v: lhs_g(v), b=b.to(cuda), x=xg, precond=lambda v: v, max_iter=20,
             tolerance=1e-3, stop='max_gain')
    assert optim.cg.last.n_iter == OO.cg.last_n_iter
    assert U.rel_l2(xg, xo) < U.REL_TOL


@pytest.mark.parametrize('name', U.GOLDEN_NAMES)
def test_update_admm_vs_golden_and_oracle(cuda, name):
    from unires_b200 import _update
    g, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    C = len(x)
    z, w = _update._admm_aux(y, sett)
    tmp = torch.zeros(y[0].dim, device=cuda)
    n_it = recipe['admm_iters']
    obj = torch.zeros(n_it, 3, dtype=torch.float64, device=cuda)
    rho = sc.rho.to(cuda)
    for it in range(n_it):
        y, z, w, tmp, obj = _update._update_admm(x, y, z, w, rho, tmp, obj, it, sett)
        iters = [i.n_iter for i in _update._update_admm.last_cg]
        assert iters == g['cg_iters'][it].tolist(), (name, it)
        for c in range(C):
            assert U.rel_l2(y[c].dat, g['y%d_it%d' % (c, it)]) < U.REL_TOL, (name, it, c)
        assert U.rel_l2(tmp, g['jtv_it%d' % it]) < 1e-3  # shrink factor: quotient of small numbers
        zs = z.flatten()[::gen_golden.SAMPLE_STRIDE]
        ws = w.flatten()[::gen_golden.SAMPLE_STRIDE]
        assert U.rel_l2(zs, g['z_sample_it%d' % it]) < 1e-3
        assert U.rel_l2(ws, g['w_sample_it%d' % it]) < 1e-3
        assert abs(z.double().norm().item() - float(g['z_norm_it%d' % it])) < 1e-3 * float(g['z_norm_it%d' % it]) + 1e-9
    assert np.allclose(obj.cpu().numpy(), g['obj'], rtol=1e-4)


def test_channel_streams_do_not_change_results(cuda):
    """The per-channel CG solves spread over several CUDA streams (sett.channel_streams) give
    bit-identical iterates and trip counts to the sequential loop of unires/_update.py:122."""
    from unires_b200 import _update
    _, recipe = U.load_golden('sr3_thick_xyz')
    res = {}
    for ns in (1, 3):
        sc = U.build(recipe, *U.port_namespaces())
        x, y, sett = U.to_device(sc, cuda)
        sett.channel_streams = ns
        z, w = _update._admm_aux(y, sett)
        tmp = torch.zeros(y[0].dim, device=cuda)
        obj = torch.zeros(2, 3, dtype=torch.float64, device=cuda)
        for it in range(2):
            y, z, w, tmp, obj = _update._update_admm(x, y, z, w, sc.rho.to(cuda), tmp, obj, it, sett)
        res[ns] = ([yc.dat.clone() for yc in y], z.clone(), w.clone(), obj.clone(),
                   [i.n_iter for i in _update._update_admm.last_cg])
    assert res[1][4] == res[3][4]
    for a, b in zip(res[1][0], res[3][0]):
        assert torch.equal(a, b)
    assert torch.equal(res[1][1], res[3][1]) and torch.equal(res[1][2], res[3][2])
    assert torch.equal(res[1][3], res[3][3])


def test_compute_nll_vs_oracle(cuda):
    from unires_b200 import _update
    _, recipe = U.load_golden('sr2_rigid')
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    got = [v.item() for v in _update._compute_nll(x, y, sett, sc.rho)]
    want = [v.item() for v in P.compute_nll(sc.x, sc.y, sc.sett, sc.rho)]
    assert np.allclose(got, want, rtol=1e-5)


def test_relaxation_alpha(cuda):
    """alpha != 1 (over-relaxation, unires/_update.py:163-164,169-170,177-178,189-190)."""
    from unires_b200 import _update
    _, recipe = U.load_golden('thickz2_scl')
    sc = U.build(recipe, *U.port_namespaces())
    sc.sett.alpha = 1.5
    x, y, sett = U.to_device(sc, cuda)
    zo, wo = P.admm_aux(sc.y)
    z, w = _update._admm_aux(y, sett)
    tmp_o, tmp = torch.zeros(sc.y[0].dim), torch.zeros(y[0].dim, device=cuda)
    obj_o = torch.zeros(2, 3, dtype=torch.float64)
    obj = torch.zeros(2, 3, dtype=torch.float64, device=cuda)
    for it in range(2):
        _, zo, wo, jo, obj_o, _ = P.update_admm(sc.x, sc.y, zo, wo, sc.rho, tmp_o, obj_o, it, sc.sett)
        y, z, w, tmp, obj = _update._update_admm(x, y, z, w, sc.rho.to(cuda), tmp, obj, it, sett)
        assert U.rel_l2(z, zo) < 1e-3 and U.rel_l2(w, wo) < 1e-3
        for c in range(len(x)):
            assert U.rel_l2(y[c].dat, sc.y[c].dat) < U.REL_TOL


def test_jtv_sharded_equals_fused(cuda):
    """norm2 (per shard) + apply == the fused single-pass prox (multi-GPU composition)."""
    import ctypes as C
    from unires_b200 import _lib, _update
    torch.manual_seed(0)
    dim = (10, 12, 14)
    Cn = 3
    ys = [torch.rand(dim, device=cuda) * 100 for _ in range(Cn)]
    lam = [0.01, 0.02, 0.015]
    w0 = torch.rand((Cn, 3) + dim, device=cuda) - 0.5
    z0 = torch.rand((Cn, 3) + dim, device=cuda) - 0.5
    vx, rho = (1.0, 1.0, 1.0), 1.7
    for alpha in (1.0, 0.8):
        z1, w1, j1 = z0.clone(), w0.clone(), torch.empty(dim, device=cuda)
        _lib.check(_lib.lib.ur_jtv_prox(_update._ptr_array(ys), _lib.ptr(z1), _lib.ptr(w1), _lib.ptr(j1),
                                        Cn, _lib.farr(lam), _lib.i3(dim), _lib.f3(vx), rho, alpha,
                                        _lib.stream()))
        z2, w2, j2 = z0.clone(), w0.clone(), torch.empty(dim, device=cuda)
        field = torch.empty(dim, device=cuda)
        for k, (lo, hi) in enumerate(((0, 2), (2, 3))):  # two "ranks"
            _lib.check(_lib.lib.ur_jtv_norm2(_update._ptr_array(ys[lo:hi]), _lib.ptr(z2[lo:hi]),
                                             _lib.ptr(w2[lo:hi]), _lib.ptr(field), hi - lo,
                                             _lib.farr(lam[lo:hi]), _lib.i3(dim), _lib.f3(vx), rho,
                                             alpha, 1 if k else 0, _lib.stream()))
        for lo, hi in ((0, 2), (2, 3)):
            _lib.check(_lib.lib.ur_jtv_apply(_update._ptr_array(ys[lo:hi]), _lib.ptr(z2[lo:hi]),
                                             _lib.ptr(w2[lo:hi]), _lib.ptr(field), _lib.ptr(j2), hi - lo,
                                             _lib.farr(lam[lo:hi]), _lib.i3(dim), _lib.f3(vx), rho,
                                             alpha, _lib.stream()))
        assert torch.allclose(z1, z2, rtol=1e-5, atol=1e-6)
        assert torch.allclose(w1, w2, rtol=1e-5, atol=1e-6)
        assert torch.allclose(j1, j2, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('name', ['denoise_1ch', 'sr3_thick_xyz', 'thickz2_scl', 'sr2_rigid'])
def test_rhs_fused_vs_general_and_oracle(cuda, name):
    """b = sum tau At x - lam div(w - rho z): fused lattice kernel, general path, oracle."""
    from oracle.nitorch_shim import spatial as OS
    from unires_b200 import _project, _update
    _, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    dim, vx = _update._geometry(y)
    g = torch.Generator().manual_seed(21)
    for c in range(len(x)):
        w = torch.rand((3,) + tuple(dim), generator=g) - 0.5
        z = torch.rand((3,) + tuple(dim), generator=g) - 0.5
        kw = dict(method=sc.sett.method, do=sc.sett.do_proj)
        ref = torch.zeros(dim)
        for n, obs in enumerate(sc.x[c]):
            ref += obs.tau * P.proj('At', obs.dat, sc.x[c], sc.y[c], n=n, **kw)
        ref -= sc.y[c].lam * OS.im_divergence(w - sc.rho * z, vx=torch.tensor(vx))
        lhs = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj, rho=sc.rho, vx_y=vx)
        out = {}
        for key, op in (('fused', lhs), ('general', None)):
            b = torch.full(dim, 7.0, device=cuda)  # must be fully overwritten
            _update._rhs(x[c], y[c], z.to(cuda), w.to(cuda), sc.rho, b, sett, dim, vx, lhs=op)
            out[key] = b
            assert U.rel_l2(b, ref) < 1e-5, (name, c, key)
        assert U.rel_l2(out['fused'], out['general']) < 1e-5


@pytest.mark.parametrize('name', ['thickz2_scl', 'sr3_thick_xyz'])
def test_update_scaling_vs_golden_and_oracle(cuda, name):
    """_update_scaling on the GPU (ur_scaling_sums + ur_scale_slices) against the reference's
    fixture and the oracle port: same line-search decisions, scl and log-likelihood."""
    from unires_b200 import _update
    g = np.load(U.GOLDEN_DIR + '/scaling_update.npz', allow_pickle=False)
    scl0 = gen_golden.SCALING_CASES[name]
    sc = gen_golden.prepare_scaling(U.build(gen_golden.RECIPES[name], *U.port_namespaces()), scl0)
    x, y, sett = U.to_device(sc, cuda)
    for k in range(gen_golden.SCALING_STEPS):
        x, sll = _update._update_scaling(x, y, sett, max_niter_gn=1, num_linesearch=6)
        got = [float(o.po.scl) for xc in x for o in xc]
        want = g[name + '_scl'][k].tolist()
        assert np.allclose(got, want, rtol=2e-4, atol=2e-6), (k, got, want)
        assert abs(float(sll) - float(g[name + '_sll'][k])) < 1e-5 * float(g[name + '_sll'][k])


def test_scaling_sums_kernel(cuda):
    """The five masked even/odd sums against torch float64 on every axis."""
    from unires_b200 import _lib
    from unires_b200._lib import lib, check, ptr, i3, stream
    g = torch.Generator().manual_seed(4)
    dim = (9, 14, 11)
    xx = torch.rand(dim, generator=g)
    xx[xx < 0.2] = 0.0
    yy = torch.rand(dim, generator=g)
    out = torch.zeros(5, dtype=torch.float64, device=cuda)
    xg, yg = xx.to(cuda), yy.to(cuda)
    for axis in range(3):
        check(lib.ur_scaling_sums(ptr(xg), ptr(yg), i3(dim), axis, ptr(out), stream()))
        m = xx != 0
        want = [torch.sum(((xx - yy) ** 2)[m], dtype=torch.float64).item()]
        for fn in (lambda a, b: b * (a - b), lambda a, b: b * b):
            for which in ('odd', 'even'):
                xs, ys, ms = (P.even_odd(t, which, axis) for t in (xx, yy, m))
                want.append(torch.sum(fn(xs, ys)[ms], dtype=torch.float64).item())
        assert np.allclose(out.cpu().numpy(), want, rtol=1e-12)


@pytest.mark.parametrize('name', sorted(gen_golden.RIGID_CASES))
def test_update_rigid_vs_golden(cuda, name):
    """_update_rigid on the GPU (ur_affine_grad + ur_rigid_sums + the operator kernels) against
    the reference's trajectory: same line-search decisions, q within 1e-3 of the step size."""
    from unires_b200 import _update
    g = np.load(U.GOLDEN_DIR + '/rigid_update.npz', allow_pickle=False)
    recipe, samp, q0 = gen_golden.RIGID_CASES[name]
    sc = gen_golden.prepare_rigid(U.build(recipe, *U.port_namespaces()), q0, P.expm)
    x, y, sett = U.to_device(sc, cuda)
    sett.rigid_basis = sc.sett.rigid_basis
    for c, xc in enumerate(x):
        for n, o in enumerate(xc):
            o.rigid_q = sc.x[c][n].rigid_q.clone()
    for k in range(gen_golden.RIGID_STEPS):
        x, sll = _update._update_rigid(x, y, sett, mean_correct=(k == gen_golden.RIGID_STEPS - 1),
                                       max_niter_gn=1, num_linesearch=6, samp=samp)
        got = np.array([o.rigid_q.cpu().tolist() for xc in x for o in xc])
        want = g[name + '_q'][k]
        assert np.allclose(got, want, rtol=2e-3, atol=2e-5), (k, got, want)
        assert abs(float(sll) - float(g[name + '_sll'][k])) < 1e-4 * float(g[name + '_sll'][k])
    rig = np.stack([o.po.rigid.cpu().numpy() for xc in x for o in xc])
    assert np.allclose(rig, g[name + '_rigid'], atol=1e-4)


def test_affine_grad_vs_oracle(cuda):
    from oracle.nitorch_shim import spatial as S
    from unires_b200 import spatial
    g = torch.Generator().manual_seed(9)
    v = torch.rand((11, 13, 9), generator=g)
    mat = torch.tensor([[0.98, 0.05, -0.02, 0.7], [-0.04, 1.01, 0.03, -0.4], [0.02, -0.03, 0.97, 1.1],
                        [0, 0, 0, 1.0]])
    shape = (12, 10, 11)
    want = S.grid_grad(v[None, None], S.affine_grid(mat, shape)[None])[0, 0]
    got = spatial.affine_grad(v.to(cuda), mat, shape)
    assert U.rel_l2(got, want) < 1e-5


@pytest.mark.parametrize('name', ['sr3_thick_xyz', 'denoise_1ch', 'sr2_rigid'])
@pytest.mark.parametrize('stop', ['max_gain', 'residual'])
def test_cg_graph_replay_is_bitwise_identical(cuda, name, stop):
    """Repeated solves with the same operator and buffers (what an ADMM run does every outer
    iteration) are captured into a CUDA graph on the second call and replayed afterwards
    (ur_tune cg_graph): iterates, trip counts, objective traces and the launch accounting are
    identical to the directly enqueued solve, including the device-side early stop."""
    from unires_b200 import _lib, optim
    _, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    b, _, lhs_g, x0 = _channel_problem(sc, 0, cuda)
    bg = b.to(cuda)
    res = {}
    try:
        for graph in (0, 1):
            _lib.check(_lib.lib.ur_tune(b'cg_graph', graph))
            x = x0.clone()
            runs = []
            for rep in range(4):
                x.copy_(x0)
                l0 = _lib.lib.ur_launch_count()
                optim.cg(A=lhs_g, b=bg, x=x, max_iter=20, tolerance=1e-3, stop=stop)
                info = optim.cg.last
                runs.append((x.clone(), info.n_iter, list(info.obj),
                             _lib.lib.ur_launch_count() - l0))
            res[graph] = runs
    finally:
        _lib.check(_lib.lib.ur_tune(b'cg_graph', 1))
    ref = res[0][0]
    for graph in (0, 1):
        for xk, n, obj, nl in res[graph]:
            assert torch.equal(xk, ref[0]) and n == ref[1] and obj == ref[2] and nl == ref[3]
