"""INTEGRATION.md Level 1: the reference's UNMODIFIED files run on top of
`unires_b200.nitorch_compat` (every nitorch call of the ADMM/CG path resolves to a kernel of
this package).  The files are executed in place from /root/reference or from the git-ignored
`baseline/_ref` install (oracle/load_reference.py); skipped when neither exists."""
import sys

import pytest
import torch

from oracle import load_reference as LR
from tests import _util as U

needs_ref = pytest.mark.skipif(not LR.available(), reason='reference files not present')


@pytest.fixture()
def compat():
    import unires_b200.nitorch_compat as nc
    saved = {k: v for k, v in sys.modules.items() if k == 'nitorch' or k.startswith('nitorch.')}
    for k in saved:
        del sys.modules[k]
    assert nc.install()
    try:
        yield nc
    finally:
        nc.uninstall()
        sys.modules.update(saved)


@needs_ref
def test_reference_files_import_on_compat(compat):
    """unires/_project.py:1-3, _update.py:5-11, _util.py:2-4, run.py:6-9, _core.py:7-19."""
    ns = LR.load_by_path('_unires_on_compat_cpu')
    assert ns._update.cg is compat.core.optim.cg
    assert ns._update.grid_grad is compat.spatial.grid_grad
    assert ns._project.grid_pull is compat.spatial.grid_pull
    assert ns.run.get_gain is compat.core.optim.get_gain
    for f in ('_update_admm', '_update_rigid', '_update_scaling', '_compute_nll', '_step_size'):
        assert callable(getattr(ns._update, f))
    assert callable(ns.run.fit) and callable(ns._core._estimate_hyperpar)


def test_compat_surface_complete(compat):
    """Every nitorch name the reference imports exists (checked without the reference)."""
    import importlib
    wanted = {
        'nitorch.spatial': ['affine_grid', 'grid_pull', 'grid_push', 'identity_grid', 'voxel_size',
                            'im_gradient', 'im_divergence', 'grid_grad', 'affine_matrix_classic',
                            'affine_basis', 'max_bb'],
        'nitorch.core.kernels': ['smooth'],
        'nitorch.core.optim': ['cg', 'get_gain', 'plot_convergence'],
        'nitorch.core.math': ['round'],
        'nitorch.core._linalg_expm': ['_expm'],
        'nitorch.core.constants': ['inf'],
        'nitorch.core.utils': ['ceil_pow'],
        'nitorch.plot.volumes': ['show_slices'],
        'nitorch.io': ['map', 'savef'],
        'nitorch.tools.img_statistics': ['estimate_noise', 'estimate_fwhm'],
        'nitorch.tools.preproc': ['atlas_crop', 'affine_align', 'atlas_align', 'reset_origin'],
        'nitorch.tools._preproc_fov': ['_bb_atlas'],
        'nitorch.tools._preproc_utils': ['_mean_space'],
    }
    for mod, names in wanted.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), (mod, n)
    import nitorch.core.math as M
    assert torch.equal(M.round(torch.tensor([1.23456]), 3), torch.tensor([1.235]))
    import nitorch.core.utils as CU
    assert CU.ceil_pow(torch.tensor([181., 217., 64.]), p=2.0, l=2.0).tolist() == [256., 256., 64.]


def test_compat_expm_matches_matrix_exp(compat):
    from nitorch.core._linalg_expm import _expm
    g = torch.Generator().manual_seed(0)
    basis = torch.zeros(6, 4, 4, dtype=torch.float64)
    for i in range(3):
        basis[i, i, 3] = 1
    basis[3, 0, 1], basis[3, 1, 0] = 1, -1
    basis[4, 0, 2], basis[4, 2, 0] = 1, -1
    basis[5, 1, 2], basis[5, 2, 1] = 1, -1
    q = torch.rand(6, generator=g, dtype=torch.float64) * 0.2
    R, dR = _expm(q, basis, grad_X=True)
    assert torch.allclose(R, torch.linalg.matrix_exp(torch.einsum('k,kij->ij', q, basis)))
    eps = 1e-6
    for i in range(6):
        dq = q.clone()
        dq[i] += eps
        fd = (_expm(dq, basis) - R) / eps
        assert torch.allclose(fd, dR[i], atol=1e-5)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('name', ['sr3_thick_xyz', 'thickz2_scl', 'sr2_rigid', 'denoise_1ch'])
def test_reference_update_admm_on_compat_equals_product(cuda, compat, monkeypatch, name):
    """The reference's own `_update_admm` (unires/_update.py:105-195, unmodified, executed in
    place) with every nitorch primitive served by this package == the product's fused
    `_update_admm`: same CG trip counts, iterates within 1e-4."""
    from unires_b200 import _update
    # the reference calls torch's F.conv3d / F.conv_transpose3d directly (unires/_project.py:
    # 153-154); cuDNN would run them in TF32 by default (1e-3 relative), which is a property of
    # the torch build, not of either implementation: compare in full float32
    monkeypatch.setattr(torch.backends.cudnn, 'allow_tf32', False)
    ns = LR.load_by_path('_unires_on_compat_gpu', mods=('struct', '_util', '_project', '_update'))
    _, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    res = {}
    for who in ('reference', 'product'):
        x, y, sett = U.to_device(sc, cuda)
        sett.device = 'cuda:0'
        sett.do_print = 0
        sett.cgs_verbose = False
        z, w = _update._admm_aux(y, sett)
        tmp = torch.zeros(y[0].dim, device=cuda)
        obj = torch.zeros(2, 3, dtype=torch.float64, device=cuda)
        rho = sc.rho.to(cuda)
        trips = []
        for it in range(2):
            if who == 'reference':
                counts = []
                cg0 = ns._update.cg

                def counting_cg(*a, **k):
                    out = cg0(*a, **k)
                    counts.append(cg0.last.n_iter)
                    return out

                ns._update.cg = counting_cg
                try:
                    y, z, w, tmp, obj = ns._update._update_admm(x, y, z, w, rho, tmp, obj, it, sett)
                finally:
                    ns._update.cg = cg0
                trips.append(counts)
            else:
                y, z, w, tmp, obj = _update._update_admm(x, y, z, w, rho, tmp, obj, it, sett)
                trips.append([i.n_iter for i in _update._update_admm.last_cg])
        res[who] = ([yc.dat.clone() for yc in y], z.clone(), w.clone(), obj.clone(), trips)
    assert res['reference'][4] == res['product'][4]
    for a, b in zip(res['reference'][0], res['product'][0]):
        assert U.rel_l2(b, a) < U.REL_TOL
    assert U.rel_l2(res['product'][1], res['reference'][1]) < 1e-3
    assert U.rel_l2(res['product'][2], res['reference'][2]) < 1e-3
    assert torch.allclose(res['product'][3], res['reference'][3], rtol=1e-4)
