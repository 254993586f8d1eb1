"""Prototype: pin the CG residual (one 256^3 volume = 67 MB) in the persisting part of the 126 MB
L2 through a stream access-policy window and time the fused CG iteration.
   python scripts/r2_l2_persist_proto.py [workload]"""
import os
import sys

import torch
from cuda.bindings import runtime as rt

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _lib, _project, struct, synth, optim  # noqa: E402


def tune(k, v):
    _lib.check(_lib.lib.ur_tune(k.encode(), int(v)))


def chk(r):
    if isinstance(r, tuple):
        err, rest = r[0], r[1:]
    else:
        err, rest = r, ()
    if int(err) != 0:
        raise RuntimeError('cuda error %s' % err)
    return rest[0] if len(rest) == 1 else rest


def set_window(stream, ptr, nbytes, ratio, prop):
    attr = rt.cudaStreamAttrValue()
    attr.accessPolicyWindow.base_ptr = ptr
    attr.accessPolicyWindow.num_bytes = nbytes
    attr.accessPolicyWindow.hitRatio = ratio
    attr.accessPolicyWindow.hitProp = prop
    attr.accessPolicyWindow.missProp = rt.cudaAccessProperty.cudaAccessPropertyStreaming
    chk(rt.cudaStreamSetAttribute(stream, rt.cudaStreamAttrID.cudaLaunchAttributeAccessPolicyWindow,
                                  attr))


def main():
    dev = torch.device('cuda:0')
    workload = sys.argv[1] if len(sys.argv) > 1 else 'sr3_256'
    max_persist = chk(rt.cudaDeviceGetAttribute(rt.cudaDeviceAttr.cudaDevAttrMaxPersistingL2CacheSize, 0))
    max_window = chk(rt.cudaDeviceGetAttribute(rt.cudaDeviceAttr.cudaDevAttrMaxAccessPolicyWindowSize, 0))
    l2 = chk(rt.cudaDeviceGetAttribute(rt.cudaDeviceAttr.cudaDevAttrL2CacheSize, 0))
    print('L2 %d MB, max persisting %d MB, max window %d MB' % (l2 >> 20, max_persist >> 20,
                                                                 max_window >> 20), flush=True)
    sc = synth.make_scenario(synth.CONFIGS[workload], _project, struct, device=dev, seed=0)
    dim = tuple(sc.y[0].dim)
    n = dim[0] * dim[1] * dim[2]
    vol = (n * 4 + 255) // 256 * 256
    vx = [float(sc.cfg['vx_y'])] * 3
    iters, reps = 20, 5
    stream = torch.cuda.current_stream(dev).cuda_stream
    for graph in (0, 1):
        tune('cg_graph', graph)
        for c in range(len(sc.x)):
            op = _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method,
                                      do=sc.sett.do_proj, rho=sc.rho, vx_y=vx)
            b = op(sc.y[c].dat) + 0.01 * torch.randn(dim, device=dev)
            x0 = sc.y[c].dat.clone()
            x = x0.clone()
            optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
            ws = _lib.workspace(op.cg_bytes, dev, 'cg')
            base = ws.data_ptr()
            # workspace: [CgState | lhs ws | r | p | Ap | p2 | b_pad | x_pad | x2]
            off = {k: op.cg_bytes - (7 - i) * vol for i, k in enumerate(['r', 'p', 'Ap', 'p2'])}
            for what, persist_mb in (('none', 0), ('r', 64), ('r', max_persist >> 20),
                                     ('Ap', max_persist >> 20), ('r+p+Ap+p2', max_persist >> 20)):
                chk(rt.cudaCtxResetPersistingL2Cache())
                if what == 'none':
                    chk(rt.cudaDeviceSetLimit(rt.cudaLimit.cudaLimitPersistingL2CacheSize, 0))
                    set_window(stream, 0, 0, 0.0, rt.cudaAccessProperty.cudaAccessPropertyNormal)
                else:
                    chk(rt.cudaDeviceSetLimit(rt.cudaLimit.cudaLimitPersistingL2CacheSize,
                                              min(persist_mb << 20, max_persist)))
                    first = what.split('+')[0]
                    nb = vol * len(what.split('+'))
                    nb = min(nb, max_window)
                    ratio = min(1.0, (persist_mb << 20) / nb)
                    set_window(stream, base + off[first], nb, ratio,
                               rt.cudaAccessProperty.cudaAccessPropertyPersisting)
                x.copy_(x0)
                optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                xref = x.clone()
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    optim.cg_fused(op, b, x, iters, 0.0, _lib.UR_STOP_NONE)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / reps / iters
                print('graph %d ch%d persist %-10s %3d MB: %7.1f us/it' % (graph, c, what,
                                                                          persist_mb, us), flush=True)
            set_window(stream, 0, 0, 0.0, rt.cudaAccessProperty.cudaAccessPropertyNormal)
            del op
    chk(rt.cudaCtxResetPersistingL2Cache())


if __name__ == '__main__':
    main()
