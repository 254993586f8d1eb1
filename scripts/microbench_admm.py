"""Time the pieces of one ADMM iteration (unires/_update.py:105-195) on a workload:
right-hand side, CG solve, objective, JTV prox.  python scripts/microbench_admm.py [workload] [tol]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _lib, _project, _update, struct, synth  # noqa: E402
from unires_b200._lib import lib, check, ptr, i3, f3, stream  # noqa: E402


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device('cuda:0')
    args = [a for a in sys.argv[1:] if '=' not in a]
    for k, v in (a.split('=') for a in sys.argv[1:] if '=' in a):
        check(lib.ur_tune(k.encode(), int(v)))
    workload = args[0] if args else 'sr3_256'
    tol = float(args[1]) if len(args) > 1 else 1e-3
    sc = synth.make_scenario(synth.CONFIGS[workload], _project, struct, device=dev, seed=0)
    sett = sc.sett
    sett.cgs_tol = tol
    x, y = sc.x, sc.y
    Cn = len(x)
    dim, vx = _update._geometry(y)
    n = dim[0] * dim[1] * dim[2]
    rho = float(sc.rho)
    z, w = _update._admm_aux(y, sett)
    tmp = torch.zeros(dim, device=dev)
    obj = torch.zeros(4, 3, dtype=torch.float64, device=dev)
    y0 = [yc.dat.clone() for yc in y]
    # a couple of real iterations so that z, w are not trivial
    for it in range(2):
        _update._update_admm(x, y, z, w, sc.rho, tmp, obj, it, sett)
    print('workload %s dim %s C %d; CG trip counts %s' % (workload, dim, Cn,
          [i.n_iter for i in _update._update_admm.last_cg]))
    GB = 1e-3  # bytes / us -> GB/s factor: bytes/us/1e3

    def rhs():
        for c in range(Cn):
            lhs = _project.LhsOperator(x[c], y[c], method=sett.method, do=sett.do_proj, rho=rho, vx_y=vx)
            _update._rhs(x[c], y[c], z[c], w[c], rho, tmp, sett, dim, vx, lhs=lhs)
    us = timed(rhs)
    b = Cn * n * 4 * (6 + 1) + sum(o.dat.numel() * 4 for xc in x for o in xc)
    print('rhs (all channels)      %8.1f us  alg %6.0f MB  %6.0f GB/s' % (us, b / 1e6, b / us * GB))

    def admm():
        for c in range(Cn):
            y[c].dat.copy_(y0[c])
        _update._update_admm(x, y, z.clone(), w.clone(), sc.rho, tmp, obj, 3, sett)
    us = timed(admm, 5)
    its = sum(i.n_iter for i in _update._update_admm.last_cg)
    print('_update_admm            %8.1f us  (%d CG iterations in total, tol %g)' % (us, its, tol))

    def nll():
        row = torch.zeros(3, dtype=torch.float64, device=dev)
        _update._nll_terms(x, y, sett, row)
    us = timed(nll)
    print('objective (_compute_nll)%8.1f us' % us)

    ys = [yc.dat for yc in y]
    lam = _lib.farr([_lib.host_scalar(yc.lam) for yc in y])

    def jtv():
        check(lib.ur_jtv_prox(_update._ptr_array(ys), ptr(z), ptr(w), ptr(tmp), Cn, lam, i3(dim), f3(vx),
                              rho, 1.0, stream()))
    us = timed(jtv)
    b = (40 * Cn + 4) * n
    print('jtv prox                %8.1f us  alg %6.0f MB  %6.0f GB/s  frac %.3f' % (us, b / 1e6, b / us * GB, b / us * GB / 6650))


if __name__ == '__main__':
    main()
