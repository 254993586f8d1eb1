"""The CUDA path against the MID-SIZE fixtures generated from the reference's own files
(oracle/gen_golden.py --mid): 40 x 96 x 260 grids, i.e. 5-12 row tiles x 3 z tiles x several
column segments of the streaming kernels, lattice-aligned and rotated (notebook-scale rigid
misalignment, even/odd slice scaling).  Volumes are compared on the fixture's strided sample
and by their float64 norm."""
import numpy as np
import pytest
import torch

from oracle import gen_golden
from tests import _util as U

pytestmark = pytest.mark.gpu


def _close(g, key, t, tol):
    s, n = gen_golden.mid_sample(t.detach().cpu())
    assert U.rel_l2(s, g[key + '_s']) < tol, key
    assert abs(float(n) - float(g[key + '_n'])) < tol * float(g[key + '_n']) + 1e-12, key


@pytest.mark.parametrize('name', sorted(gen_golden.MID_RECIPES))
@pytest.mark.parametrize('lock', [1, 0])
def test_mid_size_operators_and_admm(cuda, name, lock):
    from unires_b200 import _lib, _project, _update
    g, recipe = U.load_golden(name)
    sc = U.build(recipe, *U.port_namespaces())
    x, y, sett = U.to_device(sc, cuda)
    C = len(x)
    vx = [float(sc.cfg['vx_y'])] * 3
    _lib.check(_lib.lib.ur_tune(b'fast_lock', lock))
    try:
        for c in range(C):
            vy, vxx = gen_golden.probe_inputs(sc, c)
            po = x[c][0].po
            _close(g, 'A%d' % c, _project._proj_apply('A', vy.to(cuda)[None, None], po)[0, 0], 1e-5)
            _close(g, 'At%d' % c, _project._proj_apply('At', vxx.to(cuda)[None, None], po)[0, 0], 1e-5)
            op = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj, rho=sc.rho,
                                      vx_y=vx)
            _close(g, 'lhs%d' % c, op(vy.to(cuda)), 1e-5)
        z, w = _update._admm_aux(y, sett)
        tmp = torch.zeros(y[0].dim, device=cuda)
        obj = torch.zeros(1, 3, dtype=torch.float64, device=cuda)
        y, z, w, tmp, obj = _update._update_admm(x, y, z, w, sc.rho.to(cuda), tmp, obj, 0, sett)
        assert [i.n_iter for i in _update._update_admm.last_cg] == g['cg_iters'].tolist()
        for c in range(C):
            _close(g, 'y%d' % c, y[c].dat, U.REL_TOL)
        _close(g, 'jtv', tmp, 1e-3)
        _close(g, 'z', z, 1e-3)
        _close(g, 'w', w, 1e-3)
        assert np.allclose(obj.cpu().numpy(), g['obj'], rtol=1e-4)
    finally:
        _lib.check(_lib.lib.ur_tune(b'fast_lock', 1))
