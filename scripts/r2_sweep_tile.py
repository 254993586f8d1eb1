"""Sweep the lean kernel's tile rows / segments / stencil lag on the CG solve, per channel:
   python scripts/r2_sweep_tile.py [workload] [quick]
Every point is first checked against the default decomposition (matvec bitwise, CG iterate rel-L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _lib, _project, struct, synth, optim  # noqa: E402


def tune(k, v):
    _lib.check(_lib.lib.ur_tune(k.encode(), int(v)))


def reset():
    for k in ('fast_rpt', 'fast_pfd', 'fast_to', 'fast_segs', 'fast_lag'):
        tune(k, 0)
    tune('fast_depth', 1)
    tune('fast_lock', 1)


def main():
    dev = torch.device('cuda:0')
    workload = sys.argv[1] if len(sys.argv) > 1 else 'sr3_256'
    quick = len(sys.argv) > 2
    sc = synth.make_scenario(synth.CONFIGS[workload], _project, struct, device=dev, seed=0)
    dim = tuple(sc.y[0].dim)
    vx = [float(sc.cfg['vx_y'])] * 3
    iters, reps = 20, 5
    tune('cg_graph', 0)
    for c in range(len(sc.x)):
        op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                                  rho=sc.rho, vx_y=vx)
        b = op(sc.y[c].dat) + 0.01 * torch.randn(dim, device=dev)
        x0 = sc.y[c].dat.clone()
        x = x0.clone()
        reset()
        ref_mv = op(x0).clone()
        x.copy_(x0)
        optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
        ref_x = x.clone()
        points = []
        for rpt in (1, 2):
            tos = (8, 7, 6, 5) if rpt == 1 else (16, 15, 14, 13, 12)
            for to in tos:
                for segs in ((0,) if quick else (0, 3, 4, 5, 7, 8)):
                    for lag in (0, 1):
                        for depth in (1, 2):
                            points.append((rpt, to, segs, lag, depth))
        for rpt, to, segs, lag, depth in points:
            reset()
            tune('fast_rpt', rpt); tune('fast_to', to); tune('fast_segs', segs)
            tune('fast_lag', lag); tune('fast_depth', depth)
            tag = 'ch%d rpt %d to %2d segs %d lag %d depth %d' % (c, rpt, to, segs, lag, depth)
            try:
                mv = op(x0)
                same = bool(torch.equal(mv, ref_mv))
                x.copy_(x0)
                optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                err = float((x - ref_x).norm() / ref_x.norm())
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / reps / iters
                print('%s: %7.1f us/it  matvec_bitwise %s  cg_rel %.1e  path %d' %
                      (tag, us, same, err, _lib.lib.ur_last_lhs_path()), flush=True)
            except Exception as e:
                print('%s: failed %s' % (tag, str(e)[:80]), flush=True)
    reset()


if __name__ == '__main__':
    main()
