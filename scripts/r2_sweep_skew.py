"""Does the placement of the CG volumes modulo large powers of two matter (256^3 float volumes
are 2^26 bytes)?  python scripts/r2_sweep_skew.py [workload]"""
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
    n = dim[0] * dim[1] * dim[2]
    vx = [float(sc.cfg['vx_y'])] * 3
    iters, reps = 20, 5
    tune('cg_graph', 0)
    for c in range(len(sc.x)):
        for skew in (0, 4352, 37120, 299264, 1052928, 2101504):
            for user_skew in (0, 1):
                tune('vol_skew', skew)
                _lib._ws_cache.clear()
                op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method,
                                          do=sc.sett.do_proj, rho=sc.rho, vx_y=vx)
                # b, x0, x in one buffer: contiguous (power-of-two offsets) or skewed like the ws
                pad = (skew // 4) if user_skew else 0
                pool = torch.empty(3 * (n + pad) + 64, device=dev)
                b, x0, x = [pool[i * (n + pad): i * (n + pad) + n].view(dim) for i in range(3)]
                x0.copy_(sc.y[c].dat)
                b.copy_(op(x0) + 0.01 * torch.randn(dim, device=dev))
                x.copy_(x0)
                optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / reps / iters
                print('ch%d ws skew %8d user skew %d: %7.1f us/it' % (c, skew, user_skew, us),
                      flush=True)
                del op, pool
    tune('vol_skew', 0)


if __name__ == '__main__':
    main()
