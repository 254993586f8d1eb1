"""Debug: cell vs gather adjoint on the operators of the fit_rigid scenario (initial + final)."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gen_golden
from tests import _util as U
from unires_b200 import run, _lib, _project
cuda = torch.device('cuda:0')
g = np.load(U.GOLDEN_DIR + '/fit_rigid.npz', allow_pickle=False)
recipe = json.loads(str(g['recipe']))
sc = gen_golden.prepare_fit(U.build(recipe, *U.port_namespaces()), rigid=True)
x, y, sett = U.to_device(sc, cuda)
for k in ('max_iter', 'tolerance', 'reg_scl', 'sched_num', 'clean_fov', 'scaling',
          'unified_rigid', 'rigid_mod', 'rigid_samp', 'rigid_basis'):
    setattr(sett, k, getattr(sc.sett, k))
for c in range(len(y)):
    y[c].lam0 = torch.tensor(float(sc.y[c].lam0), device=cuda)
    for n, o in enumerate(x[c]):
        o.dim = tuple(sc.x[c][n].dat.shape)
        o.tau = torch.tensor(float(sc.x[c][n].tau), device=cuda)
        o.rigid_q = sc.x[c][n].rigid_q.clone()


def check(tag):
    gen = torch.Generator().manual_seed(3)
    for c in range(len(x)):
        for n, o in enumerate(x[c]):
            po = o.po
            u = torch.rand(tuple(po.dim_x), generator=gen).to(cuda)
            v = torch.rand(tuple(po.dim_y), generator=gen).to(cuda)
            out = {}
            for cell in (0, 1, 8):
                _lib.check(_lib.lib.ur_tune(b'rot_cell', cell))
                at = _project._proj_apply('At', u[None, None], po)[0, 0].clone()
                ata = _project._proj_apply('AtA', v[None, None], po)[0, 0].clone()
                out[cell] = (at, ata)
            s = _project.proj_struct(po, sett.method)
            print(tag, 'ch', c, 'obs', n, 'dim_x', tuple(po.dim_x), 'dim_y', tuple(po.dim_y),
                  'At rel', U.rel_l2(out[1][0], out[0][0]), U.rel_l2(out[8][0], out[0][0]),
                  'AtA rel', U.rel_l2(out[1][1], out[0][1]), U.rel_l2(out[8][1], out[0][1]))
            d = (out[1][0] - out[0][0]).abs()
            if float(d.max()) > 1e-4 * float(out[0][0].abs().max()):
                idx = torch.nonzero(d > 0.5 * d.max())[:8].tolist()
                print('   worst voxels', idx, 'max diff', float(d.max()), 'rigid', po.rigid.tolist())
    _lib.check(_lib.lib.ur_tune(b'rot_cell', 1))


check('initial')
_lib.check(_lib.lib.ur_tune(b'rot_cell', 0))
run.fit(x, y, sett)
check('final')
