"""One channel observed by three orthogonal thick-slice views (the classic multi-view
super-resolution case): time the CG iteration of lhs = sum_n tau_n An'An + rho lam^2 D'D."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _lib, _project, struct, synth, optim  # noqa: E402

dev = torch.device('cuda:0')
sc = synth.make_scenario(synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'sr3_256'], _project, struct,
                         device=dev, seed=0)
dim = tuple(sc.y[0].dim)
n = dim[0] * dim[1] * dim[2]
views = [sc.x[c][0] for c in range(len(sc.x))]
for nv in (1, 2, 3):
    op = _project.LhsOperator(views[:nv], sc.y[0], method=sc.sett.method, do=True, rho=sc.rho, vx_y=[1.0] * 3)
    b = op(sc.y[0].dat) + 0.01 * torch.randn(dim, device=dev)
    x0 = sc.y[0].dat.clone()
    for stop, tol, name in ((_lib.UR_STOP_NONE, 0.0, 'no-stop'), (_lib.UR_STOP_ENERGY, 1e-30, 'energy')):
        x = x0.clone()
        optim.cg_fused(op, b, x, 20, tol, stop)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            x.copy_(x0)
            optim.cg_fused(op, b, x, 20, tol, stop)
        e1.record()
        torch.cuda.synchronize()
        print('%d view(s), %-7s rule: %7.1f us per CG iteration (lhs path %d)'
              % (nv, name, e0.elapsed_time(e1) * 1e3 / 60, _lib.lib.ur_last_lhs_path()), flush=True)
