"""Debug: objective trajectory of run.fit on the GPU vs the reference fixture."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gen_golden
from tests import _util as U
from unires_b200 import run, _update
cuda = torch.device('cuda:0')
from unires_b200 import _lib
if len(sys.argv) > 1:
    _lib.check(_lib.lib.ur_tune(b'fast_diag_residue', int(sys.argv[1])))
g = np.load(U.GOLDEN_DIR + '/fit_sr2.npz', allow_pickle=False)
recipe = json.loads(str(g['recipe']))
sc = gen_golden.prepare_fit(U.build(recipe, *U.port_namespaces()))
x, y, sett = U.to_device(sc, cuda)
for k in ('max_iter', 'tolerance', 'reg_scl', 'sched_num', 'clean_fov', 'scaling', 'unified_rigid', 'rigid_mod'):
    setattr(sett, k, getattr(sc.sett, k))
for c in range(len(y)):
    y[c].lam0 = torch.tensor(float(sc.y[c].lam0), device=cuda)
    for n, o in enumerate(x[c]):
        o.dim = tuple(sc.x[c][n].dat.shape)
        o.tau = torch.tensor(float(sc.x[c][n].tau), device=cuda)
its = []
orig = _update._update_admm
def wrapped(*a, **k):
    r = orig(*a, **k)
    its.append([i.n_iter for i in orig.last_cg])
    return r
run._update_admm = wrapped
dat_y, *_ = run.fit(x, y, sett)
obj = run.fit.last['obj'].cpu().numpy()
ref = g['obj']
print('n_iter', run.fit.last['n_iter'], int(g['n_iter']), obj.shape, ref.shape)
rel = np.abs(obj - ref) / np.abs(ref)
for i in range(len(obj)):
    print(i, its[i], ' '.join('%.6e' % v for v in obj[i]), ' rel', ' '.join('%.2e' % v for v in rel[i]))
print('keys', list(g.keys()))
print('dat rel_l2', U.rel_l2(dat_y, g['dat_y']))
