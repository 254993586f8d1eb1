"""Debug: run.fit with unified_rigid on the GPU vs the reference fixture, per adjoint variant:
   python scripts/dbg_fit_rigid.py [rot_cell values ...]"""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gen_golden
from tests import _util as U
from unires_b200 import run, _lib
cuda = torch.device('cuda:0')
g = np.load(U.GOLDEN_DIR + '/fit_rigid.npz', allow_pickle=False)
recipe = json.loads(str(g['recipe']))
res = {}
for cell in [int(a) for a in sys.argv[1:]] or [0, 1, 8]:
    _lib.check(_lib.lib.ur_tune(b'rot_cell', cell))
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
    dat_y, _, _, R, _, _ = run.fit(x, y, sett)
    last = run.fit.last
    q = np.array([o.rigid_q.cpu().tolist() for xc in x for o in xc])
    obj = last['obj'].cpu().numpy()
    res[cell] = (dat_y.clone(), q, obj)
    print('rot_cell', cell, 'n_iter', last['n_iter'], int(g['n_iter']),
          'max |q - ref|', float(np.abs(q - g['q']).max()),
          'max obj rel', float((np.abs(obj - g['obj']) / np.abs(g['obj'])).max()),
          'dat rel_l2', U.rel_l2(dat_y, g['dat_y']), flush=True)
    n = min(len(obj), len(g['obj']))
    rel = np.abs(obj[:n, 0] - g['obj'][:n, 0]) / np.abs(g['obj'][:n, 0])
    print('  obj rel per iteration:', ' '.join('%.1e' % v for v in rel))
ks = sorted(res)
for a in ks[1:]:
    print('variant', a, 'vs', ks[0], 'dat rel_l2', U.rel_l2(res[a][0], res[ks[0]][0]),
          'max |dq|', float(np.abs(res[a][1] - res[ks[0]][1]).max()))
