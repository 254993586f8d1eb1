"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there):
adjointness, symmetry and positivity of the lhs, linearity, CG energy decrease, and the
fused lattice path agreeing with the general (unfused) path of the same operator."""
import ctypes as C

import pytest
import torch

from tests import _util as U

pytestmark = pytest.mark.gpu


def _scenario(name, cuda, n_channels=None):
    from unires_b200 import synth, _project, struct
    cfg = synth.CONFIGS[name]
    if n_channels is not None:
        cfg = dict(cfg, thick=list(cfg['thick'])[:n_channels])
    return synth.make_scenario(cfg, _project, struct, device=cuda, seed=0)


@pytest.mark.parametrize('name,nch', [('sr3_256', 3), ('thickz2_256', 1), ('denoise_181', 1)])
def test_lhs_symmetry_linearity_positivity(cuda, name, nch):
    from unires_b200 import _project
    sc = _scenario(name, cuda, nch)
    vx = [float(sc.cfg['vx_y'])] * 3
    g = torch.Generator().manual_seed(1)
    dim = tuple(sc.y[0].dim)
    u = torch.rand(dim, generator=g).to(cuda)
    v = torch.rand(dim, generator=g).to(cuda)
    for c in range(len(sc.x)):
        op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                                  rho=sc.rho, vx_y=vx)
        Au, Av = op(u), op(v)
        a = torch.sum(Au * v, dtype=torch.float64).item()
        b = torch.sum(u * Av, dtype=torch.float64).item()
        assert abs(a - b) < 1e-5 * abs(a)
        assert torch.sum(Au * u, dtype=torch.float64).item() > 0
        lin = op(2 * u - 3 * v)
        assert U.rel_l2(lin, 2 * Au - 3 * Av) < 1e-5


def test_adjoint_full_size(cuda):
    from unires_b200 import _project
    sc = _scenario('sr3_256', cuda, 3)
    for c in range(3):
        po = sc.x[c][0].po
        g = torch.Generator().manual_seed(c)
        yv = torch.rand((1, 1) + tuple(po.dim_y), generator=g).to(cuda)
        xv = torch.rand((1, 1) + tuple(po.dim_x), generator=g).to(cuda)
        Ay = _project._proj_apply('A', yv, po)
        Atx = _project._proj_apply('At', xv, po)
        a = torch.sum(Ay * xv, dtype=torch.float64).item()
        b = torch.sum(Atx * yv, dtype=torch.float64).item()
        assert abs(a - b) < 1e-5 * abs(a)


def test_fused_lattice_equals_general_path(cuda):
    """AtA through the fused lhs kernel (tau=1, rho=0) == pull/conv/conv'/push kernels."""
    from unires_b200 import _project, struct
    sc = _scenario('sr3_256', cuda, 3)
    g = torch.Generator().manual_seed(7)
    v = torch.rand(tuple(sc.y[0].dim), generator=g).to(cuda)
    for c in range(3):
        po = sc.x[c][0].po
        general = _project._proj_apply('AtA', v[None, None], po)[0, 0]
        obs = struct._input(tau=1.0, po=po)
        rec = struct._output(dim=tuple(po.dim_y), lam=0.0)
        fused = _project.LhsOperator([obs], rec, rho=0.0, vx_y=[1.0] * 3)(v)
        assert U.rel_l2(fused, general) < 1e-5


def test_cg_energy_decreases_full_size(cuda):
    from unires_b200 import _update
    sc = _scenario('sr3_256', cuda, 3)
    z, w = _update._admm_aux(sc.y, sc.sett)
    tmp = torch.zeros(tuple(sc.y[0].dim), device=cuda)
    obj = torch.zeros(2, 3, dtype=torch.float64, device=cuda)
    for it in range(2):
        _update._update_admm(sc.x, sc.y, z, w, sc.rho, tmp, obj, it, sc.sett)
        for info in _update._update_admm.last_cg:
            o = info.obj
            assert 1 <= info.n_iter <= 20
            assert all(o[k + 1] <= o[k] + 1e-9 * abs(o[k]) for k in range(len(o) - 1))
    o = obj.cpu()
    assert torch.isfinite(o).all() and o[1, 0] < o[0, 0]
