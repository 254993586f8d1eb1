"""L2 eviction-priority hints on the fused CG iteration (ur_tune l2_hints: bit 0 = once-per-launch
vectors evict_first, bit 1 = residual evict_last) x residual update direction:
   python scripts/r2_l2_hints.py [workload]"""
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
    iters, reps = 20, 5
    for c in range(len(sc.x)):
        op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                                  rho=sc.rho, vx_y=vx)
        b = op(sc.y[c].dat) + 0.01 * torch.randn(dim, device=dev)
        x0 = sc.y[c].dat.clone()
        x = x0.clone()
        ref = None
        for hints in (0, 1, 2, 3):
            for rev in (0, 1):
                tune('l2_hints', hints)
                tune('r_reverse', rev)
                x.copy_(x0)
                optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                if ref is None:
                    ref = x.clone()
                same = bool(torch.equal(x, ref))
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / reps / iters
                print('ch%d l2_hints %d r_reverse %d: %7.1f us/it  bitwise %s' %
                      (c, hints, rev, us, same), flush=True)
    tune('l2_hints', 0)
    tune('r_reverse', 0)


if __name__ == '__main__':
    main()
