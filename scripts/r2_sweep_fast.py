"""Sweep the lean kernel's knobs on the CG solve (fused COMBINE iterations), per channel:
   python scripts/r2_sweep_fast.py [workload]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _lib, _project, struct, synth, optim  # noqa: E402


def tune(k, v):
    _lib.check(_lib.lib.ur_tune(k.encode(), int(v)))


def main():
    dev = torch.device('cuda:0')
    workload = sys.argv[1] if len(sys.argv) > 1 else 'sr3_256'
    sc = synth.make_scenario(synth.CONFIGS[workload], _project, struct, device=dev, seed=0)
    dim = tuple(sc.y[0].dim)
    vx = [float(sc.cfg['vx_y'])] * 3
    iters, reps = 20, 4
    tune('cg_graph', 0)
    for c in range(len(sc.x)):
        op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                                  rho=sc.rho, vx_y=vx)
        b = op(sc.y[c].dat) + 0.01 * torch.randn(dim, device=dev)
        x0 = sc.y[c].dat.clone()
        x = x0.clone()
        for rpt in (0, 1, 2):
            for depth in (1, 2, 3):
                for pfd in (0, 2):
                    for lock in (1, 0):
                        tune('fast_rpt', rpt); tune('fast_depth', depth); tune('fast_pfd', pfd)
                        tune('fast_lock', lock)
                        try:
                            optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                            torch.cuda.synchronize()
                            e0 = torch.cuda.Event(enable_timing=True)
                            e1 = torch.cuda.Event(enable_timing=True)
                            e0.record()
                            for _ in range(reps):
                                x.copy_(x0)
                                optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                            e1.record()
                            torch.cuda.synchronize()
                            us = e0.elapsed_time(e1) * 1e3 / reps / iters
                            print('ch%d rpt %d depth %d pfd %d lock %d: %7.1f us/it' %
                                  (c, rpt, depth, pfd, lock, us), flush=True)
                        except Exception as e:
                            print('ch%d rpt %d depth %d pfd %d lock %d: failed %s' %
                                  (c, rpt, depth, pfd, lock, str(e)[:60]), flush=True)
    for k in ('fast_rpt', 'fast_pfd'):
        tune(k, 0)
    tune('fast_depth', 1)
    tune('fast_lock', 1)


if __name__ == '__main__':
    main()
