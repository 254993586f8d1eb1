"""The travelling oracle (oracle/unires_port.py) against the reference's OWN files
(/root/reference/unires/*.py imported by path on the shim).  Only runs where the
reference exists (the build container); skipped on the GPU box."""
import pytest
import torch

from oracle import load_reference as LR
from oracle import unires_port as P
from tests import _util as U

pytestmark = pytest.mark.skipif(not LR.available(), reason='/root/reference not present')


@pytest.mark.parametrize('name', U.GOLDEN_NAMES)
def test_port_matches_reference_bitwise(name):
    from oracle.adapters import reference_namespaces
    ref = LR.load_reference()
    _, recipe = U.load_golden(name)
    sp = U.build(recipe, *U.port_namespaces())
    sr = U.build(recipe, *reference_namespaces())
    for c in range(len(sp.x)):
        assert torch.equal(sp.x[c][0].dat, sr.x[c][0].dat)
        assert torch.equal(sp.y[c].dat, sr.y[c].dat)
        if sp.sett.do_proj:
            a, b = sp.x[c][0].po, sr.x[c][0].po
            assert a.dim_yx == b.dim_yx and a.ratio == b.ratio
            assert torch.equal(a.smo_ker, b.smo_ker) and torch.equal(a.mat_yx, b.mat_yx)
            assert int(a.dim_thick) == int(b.dim_thick)
    zp, wp = P.admm_aux(sp.y)
    zr, wr = ref._update._admm_aux(sr.y, sr.sett)
    tp, tr = torch.zeros(sp.y[0].dim), torch.zeros(sr.y[0].dim)
    op, orf = torch.zeros(2, 3, dtype=torch.float64), torch.zeros(2, 3, dtype=torch.float64)
    for it in range(2):
        _, zp, wp, jp, op, _ = P.update_admm(sp.x, sp.y, zp, wp, sp.rho, tp, op, it, sp.sett)
        _, zr, wr, tr, orf = ref._update._update_admm(sr.x, sr.y, zr, wr, sr.rho, tr, orf, it, sr.sett)
        for c in range(len(sp.x)):
            assert torch.equal(sp.y[c].dat, sr.y[c].dat)
        assert torch.equal(zp, zr) and torch.equal(wp, wr) and torch.equal(jp, tr)
    assert torch.equal(op, orf)


def test_step_size_matches_reference():
    from oracle.adapters import reference_namespaces
    ref = LR.load_reference()
    _, recipe = U.load_golden('sr3_thick_xyz')
    sr = U.build(recipe, *reference_namespaces())
    sr.sett.rho = None
    assert torch.equal(P.step_size(sr.x, sr.y, sr.sett), ref._update._step_size(sr.x, sr.y, sr.sett))


@pytest.mark.parametrize('name', ['thickz2_scl', 'sr3_thick_xyz'])
def test_port_update_scaling_matches_reference_bitwise(name):
    """Even/odd slice-scaling Gauss-Newton update (unires/_update.py:270-393)."""
    from oracle import gen_golden
    from oracle.adapters import reference_namespaces
    ref = LR.load_reference()
    scl0 = gen_golden.SCALING_CASES[name]
    sp = gen_golden.prepare_scaling(U.build(gen_golden.RECIPES[name], *U.port_namespaces()), scl0)
    sr = gen_golden.prepare_scaling(gen_golden.prepare_fit(
        U.build(gen_golden.RECIPES[name], *reference_namespaces()), reference=True), scl0)
    for _ in range(2):
        _, sll_p = P.update_scaling(sp.x, sp.y, sp.sett, max_niter_gn=1, num_linesearch=6)
        _, sll_r = ref._update._update_scaling(sr.x, sr.y, sr.sett, max_niter_gn=1, num_linesearch=6,
                                               verbose=0)
        assert float(sll_p) == float(sll_r)
        for xp, xr in zip(sp.x, sr.x):
            assert float(xp[0].po.scl) == float(xr[0].po.scl)


@pytest.mark.parametrize('name', ['sr2_lattice', 'thickz2_samp2'])
def test_port_update_rigid_matches_reference_bitwise(name):
    """Rigid Gauss-Newton update (unires/_update.py:198-267, 448-710) incl. sub-sampling and
    mean correction."""
    from oracle import gen_golden
    from oracle.adapters import reference_namespaces
    ref = LR.load_reference()
    recipe, samp, q0 = gen_golden.RIGID_CASES[name]
    sp = gen_golden.prepare_rigid(U.build(recipe, *U.port_namespaces()), q0, P.expm)
    sr = gen_golden.prepare_rigid(gen_golden.prepare_fit(
        U.build(recipe, *reference_namespaces()), reference=True), q0, ref._update._expm)
    for k in range(2):
        _, sll_p = P.update_rigid(sp.x, sp.y, sp.sett, mean_correct=(k == 1), max_niter_gn=1,
                                  num_linesearch=6, samp=samp)
        _, sll_r = ref._update._update_rigid(sr.x, sr.y, sr.sett, mean_correct=(k == 1),
                                             max_niter_gn=1, num_linesearch=6, verbose=0, samp=samp)
        assert float(sll_p) == float(sll_r)
        for xp, xr in zip(sp.x, sr.x):
            assert torch.equal(xp[0].rigid_q, xr[0].rigid_q)
            assert torch.equal(xp[0].po.rigid, xr[0].po.rigid)


def test_port_init_y_dat_matches_reference_bitwise():
    """Initial estimate by trilinear pull + averaging (unires/_core.py:371-399)."""
    from oracle import gen_golden
    from oracle.adapters import reference_namespaces
    ref = LR.load_reference()
    recipe = gen_golden.RECIPES['sr2_rigid']
    sp = U.build(recipe, *U.port_namespaces())
    sr = U.build(recipe, *reference_namespaces())
    sr.sett.device = 'cpu'
    P.init_y_dat(sp.x, sp.y, sp.sett)
    ref._core._init_y_dat(sr.x, sr.y, sr.sett)
    for a, b in zip(sp.y, sr.y):
        assert torch.equal(a.dat, b.dat)


def test_port_estimate_hyperpar_matches_reference_bitwise():
    """The reference's own `_estimate_hyperpar` (unires/_core.py:96-142) driving the restated
    `estimate_noise` against the port's restatement of that control flow: the `dat >= 0`
    selection, float32 casts, tau = 1 / sd^2 and mu = |foreground - noise class|."""
    from oracle import gen_golden
    from oracle.adapters import reference_namespaces
    ref = LR.load_reference()
    recipe = dict(gen_golden.RECIPES['sr3_thick_xyz'])
    sp = U.build(recipe, *U.port_namespaces())
    sr = U.build(recipe, *reference_namespaces())
    g = torch.Generator().manual_seed(9)
    for xp, xr in zip(sp.x, sr.x):  # noise everywhere (negative voxels appear), like the notebooks
        n = 25 * torch.randn(xp[0].dat.shape, generator=g)
        xp[0].dat = xp[0].dat + n
        xr[0].dat = xr[0].dat + n
    sr.x[1][0].ct = sp.x[1][0].ct = True  # one observation through the Gaussian branch
    sr.sett.do_print = 0
    sr.sett.show_hyperpar = False
    ref._core._estimate_hyperpar(sr.x, sr.sett)
    P.estimate_hyperpar(sp.x)
    for xp, xr in zip(sp.x, sr.x):
        for k in ('sd', 'tau', 'mu'):
            a, b = getattr(xp[0], k), getattr(xr[0], k)
            assert a.dtype == b.dtype == torch.float32 and torch.equal(a, b), k
        assert float(xp[0].sd) > 0
