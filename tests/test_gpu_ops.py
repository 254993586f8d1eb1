"""CUDA operators (through the C ABI) against the CPU oracle on the same seeded inputs."""
import pytest
import torch

from oracle.nitorch_shim import spatial as OS
from oracle import unires_port as P
from tests import _util as U

pytestmark = pytest.mark.gpu

DIMS = [(9, 11, 13), (16, 12, 20), (5, 7, 64)]
VX = [(1.0, 1.0, 1.0), (0.5, 0.8, 2.0)]


def _rand(shape, seed):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed)) - 0.3


@pytest.mark.parametrize('dim', DIMS)
@pytest.mark.parametrize('vx', VX)
def test_gradient_divergence_dtd(cuda, dim, vx):
    from unires_b200 import spatial, _project
    u, v = _rand(dim, 1), _rand((3,) + dim, 2)
    tvx = torch.tensor(vx)
    g = spatial.im_gradient(u.to(cuda), vx=tvx)
    assert U.rel_l2(g, OS.im_gradient(u, tvx)) < 1e-6
    d = spatial.im_divergence(v.to(cuda), vx=tvx)
    assert U.rel_l2(d, OS.im_divergence(v, tvx)) < 1e-6
    dd = _project._DtD(u.to(cuda), tvx)
    assert U.rel_l2(dd, P.dtd(u, tvx)) < 1e-6
    # composed == fused
    assert U.rel_l2(dd, spatial.im_divergence(g, vx=tvx)) < 1e-6


def _affine(seed):
    from unires_b200 import synth
    g = torch.Generator().manual_seed(seed)
    t = (torch.rand(3, generator=g) * 4 - 2).tolist()
    r = (torch.rand(3, generator=g) * 0.2 - 0.1).tolist()
    m = synth.rigid_matrix(t, r)
    m[:3, :3] *= 0.9
    return m.float()


@pytest.mark.parametrize('order', [1, 0])
@pytest.mark.parametrize('lazy', [True, False])
def test_pull_push_vs_oracle(cuda, order, lazy):
    from unires_b200 import spatial
    src_dim, out_dim = (12, 10, 14), (9, 13, 11)
    mat = _affine(3)
    src, val = _rand((1, 1) + src_dim, 4), _rand((1, 1) + out_dim, 5)
    ogrid = OS.affine_grid(mat, out_dim)[None]
    grid = spatial.affine_grid(mat.to(cuda), out_dim)[None, ...]
    if not lazy:
        grid = grid.materialize()
        assert U.rel_l2(grid, ogrid) < 1e-6
    pulled = spatial.grid_pull(src.to(cuda), grid, interpolation=order, bound='zero', extrapolate=False)
    ref = OS.grid_pull(src, ogrid, interpolation=order)
    assert pulled.shape == ref.shape and U.rel_l2(pulled, ref) < 1e-5
    pushed = spatial.grid_push(val.to(cuda), grid, shape=src_dim, interpolation=order)
    ref = OS.grid_push(val, ogrid, shape=src_dim, interpolation=order)
    assert pushed.shape == ref.shape and U.rel_l2(pushed, ref) < 1e-5


def test_pull_integer_shift_is_exact(cuda):
    from unires_b200 import spatial
    src = _rand((1, 1, 6, 7, 8), 6)
    m = torch.eye(4)
    m[:3, 3] = torch.tensor([1.0, -2.0, 3.0])
    out = spatial.grid_pull(src.to(cuda), spatial.affine_grid(m.to(cuda), (6, 7, 8))[None, ...])
    ref = OS.grid_pull(src, OS.affine_grid(m, (6, 7, 8))[None])
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize('name', [n for n in U.GOLDEN_NAMES if n != 'denoise_1ch'])
def test_proj_apply_vs_oracle_and_golden(cuda, name):
    from oracle import gen_golden
    from unires_b200 import _project
    g, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    for c in range(len(x)):
        vy, vx = gen_golden.probe_inputs(sc, c)
        po_o, po = sc.x[c][0].po, x[c][0].po
        for op, v in (('A', vy), ('At', vx), ('AtA', vy)):
            out = _project._proj_apply(op, v.to(cuda)[None, None], po, method=sett.method)[0, 0]
            ref = P.proj_apply(op, v[None, None], po_o, method=sett.method)[0, 0]
            assert out.shape == ref.shape
            assert U.rel_l2(out, ref) < 1e-5, (name, op, c)
            assert U.rel_l2(out, g['%s%d' % (op, c)]) < 1e-5, (name, op, c, 'golden')


@pytest.mark.parametrize('name', U.GOLDEN_NAMES)
def test_lhs_vs_oracle_and_golden(cuda, name):
    from oracle import gen_golden
    from unires_b200 import _project
    g, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    vx_y = torch.ones(3) * float(sc.cfg['vx_y'])
    for c in range(len(x)):
        vy, _ = gen_golden.probe_inputs(sc, c)
        out = _project._proj('AtA', vy.to(cuda), x[c], y[c], method=sett.method, do=sett.do_proj,
                             rho=sc.rho, vx_y=vx_y)
        ref = P.proj('AtA', vy, sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                     rho=sc.rho, vx_y=vx_y)
        assert U.rel_l2(out, ref) < 1e-5, (name, c)
        assert U.rel_l2(out, g['lhs%d' % c]) < 1e-5, (name, c, 'golden')
        # fused dot-product epilogue
        op = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj, rho=sc.rho,
                                  vx_y=vx_y)
        dot = torch.zeros(1, dtype=torch.float64, device=cuda)
        out2 = op(vy.to(cuda), dot=dot)
        want = torch.sum(vy * ref, dtype=torch.float64).item()
        assert abs(dot.item() - want) < 1e-6 * abs(want)
        # the general (rotated) path pushes with atomics: summation order varies run to run
        assert U.rel_l2(out, out2) < 1e-6


def test_lhs_two_observations_and_odd_dims(cuda):
    """Two repeats of one channel with different thick axes, odd extents (no 16-byte rows)."""
    from unires_b200 import _project, struct
    dim_y = (13, 15, 17)
    mat_y = torch.eye(4, dtype=torch.float64)
    obs_o, obs_g = [], []
    for axis, f in ((0, 2), (2, 3)):
        scl = [1.0, 1.0, 1.0]
        scl[axis] = float(f)
        mat_x = torch.diag(torch.tensor(scl + [1.0], dtype=torch.float64))
        dim_x = tuple(int(d // s) for d, s in zip(dim_y, scl))
        po_o = P.proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=2, prof_tp=0, scl=0.07)
        po_g = _project._proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=2, prof_tp=0, scl=0.07,
                                   device=cuda)
        tau = 0.01 * (1 + axis)
        obs_o.append(P.Observation(torch.zeros(dim_x), mat_x, tau=tau, po=po_o))
        obs_g.append(struct._input(dat=None, tau=tau, po=po_g))
    rec_o = P.Recon(torch.zeros(dim_y), mat_y, lam=0.2)
    rec_g = struct._output(dat=None, dim=dim_y, mat=mat_y, lam=0.2)
    v = _rand(dim_y, 9)
    vx = torch.ones(3)
    out = _project._proj('AtA', v.to(cuda), obs_g, rec_g, rho=2.0, vx_y=vx)
    ref = P.proj('AtA', v, obs_o, rec_o, rho=2.0, vx_y=vx)
    assert U.rel_l2(out, ref) < 1e-5


def test_apply_scaling_and_conv_axis(cuda):
    import ctypes as C
    from unires_b200 import _project, _lib
    from torch.nn import functional as F
    v = _rand((1, 1, 6, 9, 10), 11)
    for dim in range(3):
        out = _project._apply_scaling(v.to(cuda), 0.3, dim)
        assert U.rel_l2(out, P.apply_scaling(v, torch.tensor(0.3), dim)) < 1e-6
    ker = [0.1, 0.2, 0.4, 0.2, 0.1]
    for axis, stride in ((0, 1), (1, 2), (2, 3)):
        shape = [1, 1, 1, 1, 1]
        shape[2 + axis] = 5
        k = torch.tensor(ker).reshape(shape)
        st = [1, 1, 1]
        st[axis] = stride
        ref = F.conv3d(v, k, stride=st)
        out = torch.empty(ref.shape[2:], device=cuda)
        _lib.check(_lib.lib.ur_conv_axis(_lib.ptr(v.to(cuda)), _lib.i3(v.shape[2:]), _lib.ptr(out),
                                         axis, _lib.farr(ker), 5, stride, 0, _lib.stream()))
        assert U.rel_l2(out, ref[0, 0]) < 1e-6
        back = F.conv_transpose3d(ref, k, stride=st)
        out2 = torch.empty(back.shape[2:], device=cuda)
        _lib.check(_lib.lib.ur_conv_axis(_lib.ptr(out), _lib.i3(out.shape), _lib.ptr(out2), axis,
                                         _lib.farr(ker), 5, stride, 1, _lib.stream()))
        assert U.rel_l2(out2, back[0, 0]) < 1e-5


def test_check_adjoint(cuda):
    from unires_b200 import _project, synth
    cfg = synth.scaled(synth.CONFIGS['sr3_256'], (40, 44, 36))
    for c, rigid in ((0, None), (1, synth.rigid_matrix((1.5, -1.0, 0.5), (0.05, -0.03, 0.04)))):
        dim_x, mat_x, dim_y, mat_y = synth.geometry(cfg, c)
        po = _project._proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=2, prof_tp=0,
                                 scl=0.1, device=cuda)
        val = _project._check_adjoint(po, 'super-resolution', 'zero', 'linear')
        assert abs(val.item()) < 1e-2  # float32 sums of O(1e4) terms (reference prints ~1e-6..1e-3)


def test_init_y_dat_vs_oracle(cuda):
    """Initial estimate (unires/_core.py:371-399) on the CUDA pull kernel against the port."""
    from oracle import gen_golden
    from unires_b200 import io
    sc = U.build(gen_golden.RECIPES['sr2_rigid'], *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    P.init_y_dat(sc.x, sc.y, sc.sett)
    io._init_y_dat(x, y, sett)
    for a, b in zip(y, sc.y):
        assert U.rel_l2(a.dat, b.dat) < 1e-5


def test_rotated_adjoint_is_deterministic_and_matches_scatter(cuda):
    """The adjoint of a ROTATED operator runs as a gather (one thread per recon voxel, no
    atomics): bit-identical from run to run, and equal to the atomic scatter form
    (ur_affine_push) up to summation order."""
    import ctypes as C
    from oracle import gen_golden
    from unires_b200 import _lib, _project
    from unires_b200._lib import lib, check, ptr, i3, stream
    sc = U.build(gen_golden.RECIPES['sr2_rigid'], *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    po = x[0][0].po
    g = torch.Generator().manual_seed(2)
    v = torch.rand(tuple(po.dim_x), generator=g).to(cuda)
    outs = [_project._proj_apply('At', v[None, None], po, method=sett.method)[0, 0].clone()
            for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    want = P.proj_apply('At', v.cpu()[None, None], sc.x[0][0].po, method=sc.sett.method)[0, 0]
    assert U.rel_l2(outs[0], want) < 1e-5
    # pure push: gather (inside the operator, denoising method = pull/push only) vs scatter API
    s = _project.proj_struct(po, 'denoising')
    w = torch.rand(tuple(po.dim_x), generator=g).to(cuda)
    gather = _project._proj_apply('At', w[None, None], po, method='denoising')[0, 0]
    scatter = torch.zeros(tuple(po.dim_y), device=cuda)
    check(lib.ur_affine_push(ptr(w), i3(po.dim_x), s.mat, ptr(scatter), i3(po.dim_y), 1, 0, 1.0,
                             stream()))
    assert U.rel_l2(gather, scatter) < 1e-6


@pytest.mark.parametrize('name', ['sr2_rigid', 'mid_sr3_rigid'])
def test_rotated_fused_kernels_equal_general_path(cuda, name):
    """Rotated operators: the in-tile forward kernel + gather adjoint (csrc/rot.cuh: pull, slice
    profile, scaling and the transposed profile in shared memory; adjoint gathered inside the
    lhs kernel) against the pull / conv / scale / conv' / push chain they replace
    (ur_tune rot_fused=0), for A, At, AtA and the CG left-hand side; both are deterministic."""
    from oracle import gen_golden
    from unires_b200 import _lib, _project
    _, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    vx_y = [float(sc.cfg['vx_y'])] * 3
    res = {}
    try:
        for fused in (1, 0):
            _lib.check(_lib.lib.ur_tune(b'rot_fused', fused))
            outs = []
            for c in range(len(x)):
                vy, vx = gen_golden.probe_inputs(sc, c)
                po = x[c][0].po
                for op, v in (('A', vy), ('At', vx), ('AtA', vy)):
                    outs.append(_project._proj_apply(op, v.to(cuda)[None, None], po)[0, 0].clone())
                lhs = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj,
                                           rho=sc.rho, vx_y=vx_y)
                a = lhs(vy.to(cuda))
                b = lhs(vy.to(cuda))
                assert torch.equal(a, b)  # no atomics: bit-reproducible
                outs.append(a)
            res[fused] = outs
    finally:
        _lib.check(_lib.lib.ur_tune(b'rot_fused', 1))
    for a, b in zip(res[1], res[0]):
        assert a.shape == b.shape and U.rel_l2(a, b) < 2e-6


@pytest.mark.parametrize('name', ['sr2_rigid', 'mid_sr3_rigid'])
def test_rotated_cell_adjoint_equals_gather(cuda, name):
    """Adjoint pull of rotated operators through per-cell corner coefficients
    (rot_adjoint_cell_kernel: scatter into shared-memory cells in colour passes, then a gather
    per voxel) against the per-voxel candidate gather (ur_tune rot_cell=0), with two and with
    eight colour passes, for At, AtA and the CG left-hand side; every variant bit-reproducible."""
    from oracle import gen_golden
    from unires_b200 import _lib, _project
    _, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    vx_y = [float(sc.cfg['vx_y'])] * 3
    res = {}
    try:
        for cell in (1, 8, 0):
            _lib.check(_lib.lib.ur_tune(b'rot_cell', cell))
            outs = []
            for c in range(len(x)):
                vy, vx = gen_golden.probe_inputs(sc, c)
                po = x[c][0].po
                for op, v in (('At', vx), ('AtA', vy)):
                    a = _project._proj_apply(op, v.to(cuda)[None, None], po)[0, 0].clone()
                    b = _project._proj_apply(op, v.to(cuda)[None, None], po)[0, 0]
                    assert torch.equal(a, b)
                    outs.append(a)
                lhs = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj,
                                           rho=sc.rho, vx_y=vx_y)
                a = lhs(vy.to(cuda)).clone()
                assert torch.equal(a, lhs(vy.to(cuda)))
                outs.append(a)
            res[cell] = outs
    finally:
        _lib.check(_lib.lib.ur_tune(b'rot_cell', 1))
    for a, b, c in zip(res[1], res[8], res[0]):
        assert U.rel_l2(a, c) < 2e-6 and U.rel_l2(b, c) < 2e-6


@pytest.mark.parametrize('rot', [(0.0, 0.0, 0.7854), (0.5, -0.4, 0.7), (1e-5, 0.0, -2e-5),
                                 (0.1, -0.1, 0.1)])
def test_rotated_cell_adjoint_any_rotation(cuda, rot):
    """Large, tiny and notebook-sized rotations: whatever path the operator selects (two or
    eight colours, or the gather when the tile's pre-image is too large) equals the gather, and
    P' is the exact transpose of P (<P v, u> = <v, P' u> in float64 sums)."""
    from unires_b200 import _lib, _project, synth
    cfg = synth.scaled(synth.CONFIGS['sr3_256'], (48, 56, 60))
    dim_x, mat_x, dim_y, mat_y = synth.geometry(cfg, 2)
    rigid = synth.rigid_matrix((1.5, -2.0, 0.75), rot)
    po = _project._proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=2, prof_tp=0,
                             scl=0.0, device=cuda)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(tuple(po.dim_x), generator=g).to(cuda)
    v = torch.rand(tuple(po.dim_y), generator=g).to(cuda)
    outs = {}
    try:
        for cell in (1, 0):
            _lib.check(_lib.lib.ur_tune(b'rot_cell', cell))
            outs[cell] = _project._proj_apply('At', u[None, None], po)[0, 0].clone()
    finally:
        _lib.check(_lib.lib.ur_tune(b'rot_cell', 1))
    assert U.rel_l2(outs[1], outs[0]) < 2e-6
    Av = _project._proj_apply('A', v[None, None], po)[0, 0]
    lhs_ = float((Av.double() * u.double()).sum())
    rhs_ = float((v.double() * outs[1].double()).sum())
    assert abs(lhs_ - rhs_) <= 1e-4 * abs(lhs_)


@pytest.mark.parametrize('dim_y,scl,shift', [((40, 36, 52), (2, 2, 2), (0, 0, 0)),
                                             ((33, 30, 44), (2, 3, 1), (1, -2, 0)),
                                             ((26, 41, 36), (1, 2, 4), (0, 3, -1)),
                                             ((24, 24, 24), (3, 2, 2), (0, 0, 0))])
def test_multi_axis_lattice_through_low_res_image(cuda, dim_y, scl, shift):
    """Lattice operators decimated along several axes (csrc/lattice_nd.cu: nd_down + nd_up, the
    low-resolution image is the only intermediate) against the oracle's pull / dense conv3d /
    conv_transpose3d / push (unires/_project.py:147-179), for A, At, AtA and the CG left-hand
    side with its dot product; also against the chained single-axis passes they replace."""
    from oracle.adapters import port_ops
    from unires_b200 import _lib, _project, struct
    mat_y = torch.eye(4, dtype=torch.float64)
    tr = torch.eye(4, dtype=torch.float64)
    tr[:3, 3] = torch.tensor([float(s) for s in shift], dtype=torch.float64)
    mat_x = mat_y @ tr @ torch.diag(torch.tensor([float(s) for s in scl] + [1.0], dtype=torch.float64))
    dim_x = tuple(int((d - abs(t)) // s) - 1 for d, s, t in zip(dim_y, scl, shift))
    po_o = port_ops._proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=2, prof_tp=0, gap=0.0, scl=0.0)
    po = _project._proj_info(dim_y, mat_y, dim_x, mat_x, prof_ip=2, prof_tp=0, gap=0.0,
                             device=cuda, scl=0.0)
    g = torch.Generator().manual_seed(3)
    vy = torch.rand(dim_y, generator=g) + 1.0
    vx = torch.rand(dim_x, generator=g)
    tau, lam, rho = 0.0016, 0.01, 4.0
    obs_o = type('O', (), {})()
    obs_o.po, obs_o.tau = po_o, torch.tensor(tau)
    rec_o = type('R', (), {})()
    rec_o.dim, rec_o.lam = dim_y, torch.tensor(lam)
    want = {op: P.proj_apply(op, v[None, None], po_o)[0, 0]
            for op, v in (('A', vy), ('At', vx), ('AtA', vy))}
    want['lhs'] = P.proj('AtA', vy, [obs_o], rec_o, rho=torch.tensor(rho), vx_y=torch.ones(3))
    obs = struct._input(tau=tau, po=po)
    rec = struct._output(dim=dim_y, lam=lam)
    res = {}
    try:
        for fused in (1, 0):
            _lib.check(_lib.lib.ur_tune(b'nd_fused', fused))
            got = {op: _project._proj_apply(op, v.to(cuda)[None, None], po)[0, 0]
                   for op, v in (('A', vy), ('At', vx), ('AtA', vy))}
            lhs = _project.LhsOperator([obs], rec, rho=rho, vx_y=[1.0] * 3)
            dot = torch.zeros(1, dtype=torch.float64, device=cuda)
            got['lhs'] = lhs(vy.to(cuda), dot=dot)
            res[fused] = got
            for k in want:
                assert U.rel_l2(got[k], want[k]) < 1e-5, (k, fused)
            d = torch.sum(vy.double() * want['lhs'].double()).item()
            assert abs(dot.item() - d) < 1e-5 * abs(d)
        assert _lib.lib.ur_last_lhs_path() != 4  # fused = 0 ran last
    finally:
        _lib.check(_lib.lib.ur_tune(b'nd_fused', 1))
    lhs = _project.LhsOperator([obs], rec, rho=rho, vx_y=[1.0] * 3)
    lhs(vy.to(cuda))
    if dim_y[2] % 4 == 0 and sum(s > 1 for s in scl) >= 2:
        assert _lib.lib.ur_last_lhs_path() == 4


@pytest.mark.parametrize('dim_y,combo', [((20, 24, 28), 'rot+rot'), ((20, 24, 28), 'rot+lattice'),
                                         ((19, 23, 27), 'rot+rot'), ((20, 24, 28), 'rot+rot+rot')])
def test_lhs_several_observations_with_rotated_operators(cuda, dim_y, combo):
    """One channel observed by several scans of which some are rigidly mis-aligned
    (unires/_project.py:80-84 sums tau_n An'An over the repeats): two rotated terms are gathered
    in the quad kernel, rotated + lattice terms and odd nz in the direct kernel, a third rotated
    term goes through the (unfused-launch) general path -- all against the oracle, and the CG
    solve keeps the oracle's trip count."""
    from oracle.nitorch_shim.core import optim as OO
    from unires_b200 import _project, optim, struct, synth
    mat_y = torch.eye(4, dtype=torch.float64)
    specs = {'rot': [((1.2, -0.8, 0.5), (0.05, -0.03, 0.08)), ((-0.7, 1.1, -0.4), (-0.06, 0.04, 0.02)),
                     ((0.4, 0.3, -0.9), (0.02, 0.07, -0.05))]}
    obs_o, obs_g = [], []
    k_rot = 0
    for n, kind in enumerate(combo.split('+')):
        axis, f = (0, 2) if n % 2 == 0 else (2, 2)
        scl = [1.0, 1.0, 1.0]
        scl[axis] = float(f)
        mat_x = torch.diag(torch.tensor(scl + [1.0], dtype=torch.float64))
        dim_x = tuple(int(d // s) for d, s in zip(dim_y, scl))
        rigid = None
        if kind == 'rot':
            rigid = synth.rigid_matrix(*specs['rot'][k_rot])
            k_rot += 1
        po_o = P.proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=2, prof_tp=0, scl=0.05)
        po_g = _project._proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=2, prof_tp=0,
                                   scl=0.05, device=cuda)
        tau = 0.01 * (1 + n)
        obs_o.append(P.Observation(torch.zeros(dim_x), mat_x, tau=tau, po=po_o))
        obs_g.append(struct._input(dat=None, tau=tau, po=po_g))
    rec_o = P.Recon(torch.zeros(dim_y), mat_y, lam=0.2)
    rec_g = struct._output(dat=None, dim=dim_y, mat=mat_y, lam=0.2)
    v = _rand(dim_y, 9) + 1.0
    vx = torch.ones(3)
    lhs_o = lambda t: P.proj('AtA', t, obs_o, rec_o, rho=2.0, vx_y=vx)
    op = _project.LhsOperator(obs_g, rec_g, rho=2.0, vx_y=vx)
    assert U.rel_l2(op(v.to(cuda)), lhs_o(v)) < 1e-5
    b = lhs_o(_rand(dim_y, 10) + 0.5)
    xo = torch.zeros(dim_y)
    OO.cg(A=lhs_o, b=b, x=xo, max_iter=12, tolerance=1e-3, stop='max_gain')
    xg = torch.zeros(dim_y, device=cuda)
    optim.cg(A=op, b=b.to(cuda), x=xg, max_iter=12, tolerance=1e-3, stop='max_gain')
    assert optim.cg.last.n_iter == OO.cg.last_n_iter
    assert U.rel_l2(xg, xo) < U.REL_TOL


def test_backproject_one_pass_equals_general_adjoint(cuda):
    """ur_backproject: out = scale * sum_n An' x_n in one pass for lattice-aligned observations
    (the device-side initial estimate of the host pipeline) against the operator's own adjoint;
    rotated observations are declined (nothing launched)."""
    from unires_b200 import _project, _update, struct, synth
    cfg = synth.scaled(synth.CONFIGS['sr3_256'], (40, 48, 56))
    g = torch.Generator().manual_seed(9)
    for c in range(3):
        dim_x, mat_x, dim_y, mat_y = synth.geometry(cfg, c)
        for rigid, scl in ((None, 0.0), (None, 0.07),
                           (synth.rigid_matrix((1.0, -0.5, 0.25), (0.03, 0.0, -0.02)), 0.0)):
            po = _project._proj_info(dim_y, mat_y, dim_x, mat_x, rigid=rigid, prof_ip=2,
                                     prof_tp=0, scl=scl, device=cuda)
            obs = struct._input(dat=torch.rand(tuple(po.dim_x), generator=g).to(cuda), tau=0.4,
                                po=po)
            rec = struct._output(dim=dim_y, mat=mat_y, lam=0.2)
            lhs = _project.LhsOperator([obs], rec, rho=1.0, vx_y=[1.0, 1.0, 1.0])
            scale = (torch.rand(dim_y, generator=g) + 0.5).to(cuda)
            out = torch.full(dim_y, float('nan'), device=cuda)
            ok = _update._backproject([obs], out, lhs, scale)
            if rigid is not None:
                assert not ok
                continue
            assert ok
            want = _project._proj_apply('At', obs.dat[None, None], po)[0, 0] * scale
            assert U.rel_l2(out, want) < 1e-6
            plain = torch.empty(dim_y, device=cuda)
            assert _update._backproject([obs], plain, lhs)
            assert U.rel_l2(plain * scale, want) < 1e-6
