"""Time the lhs matvec per channel / variant / chunk size (CUDA events, many launches)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _lib, _project, struct, synth  # noqa: E402


def tune(k, v):
    _lib.check(_lib.lib.ur_tune(k.encode(), int(v)))


def time_op(op, v, n=30):
    out = op(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ws = _lib.workspace(op.lhs_bytes, v.device, 'lhs')
    e0.record()
    for _ in range(n):
        _lib.check(_lib.lib.ur_lhs_apply(C.byref(op.c), _lib.ptr(v), _lib.ptr(out), None, _lib.ptr(ws),
                                         ws.numel(), _lib.stream()))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def main():
    dev = torch.device('cuda:0')
    workload = sys.argv[1] if len(sys.argv) > 1 else 'sr3_256'
    mcs = [int(a) for a in sys.argv[2:]] or [0]
    sc = synth.make_scenario(synth.CONFIGS[workload], _project, struct, device=dev, seed=0)
    dim = tuple(sc.y[0].dim)
    n = dim[0] * dim[1] * dim[2]
    v = torch.rand(dim, device=dev)
    # flush helper: big buffer written between measurements is not needed (n=30 launches of
    # 134 MB traffic each cycle through L2), but report both
    peak = 6551.7
    vx = [float(sc.cfg['vx_y'])] * 3
    for c in range(len(sc.x)):
        op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method, do=sc.sett.do_proj,
                                  rho=sc.rho, vx_y=vx)
        for variant, rpt in ((0, 1), (0, 2)):
            for mc in mcs:
                tune('lhs_variant', variant)
                tune('stream_mc', mc)
                tune('stream_rpt', rpt)
                us = time_op(op, v)
                print('channel %d stream rpt %d q %3d: %8.1f us  %7.1f GB/s  frac %.3f'
                      % (c, rpt, mc, us, 8 * n / us / 1e3, 8 * n / us / 1e3 / peak), flush=True)
    tune('stream_rpt', 0)
    tune('lhs_variant', 0)
    tune('stream_mc', 0)
    # denoise lhs (no projection)
    op = _project.LhsOperator([struct._input(tau=0.01)], struct._output(dim=dim, lam=0.1), do=False,
                              rho=1.0, vx_y=vx)
    for rpt in (1, 2):
        for mc in mcs:
            tune('stream_mc', mc)
            tune('stream_rpt', rpt)
            us = time_op(op, v)
            print('denoise stream rpt %d q %3d: %8.1f us  %7.1f GB/s  frac %.3f'
                  % (rpt, mc, us, 8 * n / us / 1e3, 8 * n / us / 1e3 / peak), flush=True)
    tune('stream_rpt', 0)
    tune('stream_mc', 0)
    # plain copy for reference
    a = torch.empty_like(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        a.copy_(v)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    print('torch copy: %8.1f us  %7.1f GB/s' % (us, 8 * n / us / 1e3))


if __name__ == '__main__':
    main()
