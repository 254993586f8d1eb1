"""Time the device-resident CG solve per channel (fixed trip count, tolerance 0) and the
average matvec launch inside it:  python scripts/microbench_cg.py [workload] [knob=value ...]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _lib, _project, struct, synth, optim  # noqa: E402


def tune(k, v):
    _lib.check(_lib.lib.ur_tune(k.encode(), int(v)))


def main():
    dev = torch.device('cuda:0')
    args = [a for a in sys.argv[1:] if '=' not in a]
    knobs = [a.split('=') for a in sys.argv[1:] if '=' in a]
    workload = args[0] if args else 'sr3_256'
    iters = int(args[1]) if len(args) > 1 else 20
    reps = int(args[2]) if len(args) > 2 else 5
    for k, v in knobs:
        tune(k, v)
    sc = synth.make_scenario(synth.CONFIGS[workload], _project, struct, device=dev, seed=0)
    dim = tuple(sc.y[0].dim)
    n = dim[0] * dim[1] * dim[2]
    vx = [float(sc.cfg['vx_y'])] * 3
    peak = 6650.0
    for c in range(len(sc.x)):
        op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                                  rho=sc.rho, vx_y=vx)
        b = op(sc.y[c].dat) + 0.01 * torch.randn(dim, device=dev)
        x0 = sc.y[c].dat.clone()
        x = x0.clone()
        optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
        torch.cuda.synchronize()
        _lib.lib.ur_profile_matvec(0 if os.environ.get('NOPROF') else 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            x.copy_(x0)
            optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
        e1.record()
        torch.cuda.synchronize()
        tot, cnt, bpv = C.c_double(0), C.c_int32(0), C.c_double(0)
        _lib.check(_lib.lib.ur_profile_matvec_read(C.byref(tot), C.byref(cnt), C.byref(bpv)))
        _lib.lib.ur_profile_matvec(0)
        us_it = e0.elapsed_time(e1) * 1e3 / reps / iters
        mv_us = max(tot.value * 1e3 / max(cnt.value, 1), 1e-9)
        mv_b = bpv.value / max(cnt.value, 1) * n
        print('channel %d: %7.1f us/CG-it (%6.0f it/s)  matvec %6.1f us  %6.0f GB/s  frac %.3f  '
              '[36 B/voxel iteration: %5.0f GB/s frac %.3f]'
              % (c, us_it, 1e6 / us_it, mv_us, mv_b / mv_us / 1e3, mv_b / mv_us / 1e3 / peak,
                 36.0 * n / us_it / 1e3, 36.0 * n / us_it / 1e3 / peak), flush=True)


if __name__ == '__main__':
    main()
