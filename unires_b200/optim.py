"""nitorch.core.optim drop-ins: `cg` and `get_gain`.

`cg` keeps nitorch's signature as called at unires/_update.py:142-148.  When
`A` is a :class:`unires_b200._project.LhsOperator` and the preconditioner is
the identity (precond=None -- UniRes' `lambda x: x`, unires/_update.py:137)
the whole solve -- matvecs, float64 dot products, alpha/beta, the
``|gain| < tolerance`` stop test -- runs on the device with no host
synchronisation (ur_cg_solve).  For any other callable the same CUDA vector
kernels are driven from a host loop (one sync per iteration for the stop test,
like nitorch).

Stop rule: nitorch keeps only the first letter of
`stop`; 'e' (or 'residual') selects sqrt(r.z), anything else -- including
UniRes' 'max_gain' -- the energy 0.5 x'Ax - b'x.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, ptr, stream, require_cuda_f32
from ._project import LhsOperator


def get_gain(obj, monotonicity='increasing'):
    """(obj[-1]-obj[-2]) / (max-min) (sign by monotonicity); inf for one value."""
    if len(obj) <= 1:
        return torch.tensor(float('inf'), dtype=obj.dtype, device=obj.device)
    if monotonicity == 'increasing':
        gain = obj[-1] - obj[-2]
    elif monotonicity == 'decreasing':
        gain = obj[-2] - obj[-1]
    else:
        raise ValueError('Undefined monotonicity')
    return gain / (torch.max(obj) - torch.min(obj))


def stop_rule(stop, tolerance, verbose=False):
    if not (tolerance or verbose):
        return _lib.UR_STOP_NONE
    if stop == 'residual':
        stop = 'e'
    elif stop == 'norm':
        stop = 'a'
    return _lib.UR_STOP_RESIDUAL if stop[0].lower() == 'e' else _lib.UR_STOP_ENERGY


_STATE_BYTES = 4096  # >= sizeof(CgState) in csrc/solver.cuh


class CgInfo:
    """Result handle of the last device-side solve (lazy: reading syncs)."""

    def __init__(self, ws, stream_ptr):
        self._ws, self._stream = ws, stream_ptr
        self._n, self._obj = None, None

    def _fetch(self):
        if self._n is None:
            n = C.c_int32(0)
            obj = (C.c_double * (_lib.UR_CG_MAX_ITER + 1))()
            check(lib.ur_cg_fetch(ptr(self._ws), C.byref(n), obj, _lib.UR_CG_MAX_ITER + 1,
                                  self._stream))
            self._n, self._obj = n.value, list(obj[:n.value + 1])

    @property
    def n_iter(self):
        self._fetch()
        return self._n

    @property
    def obj(self):
        self._fetch()
        return self._obj


def cg_fused(A, b, x, max_iter, tolerance, rule, variant=0, ws=None):
    """Device-resident solve of A x = b, in place on x.  Returns CgInfo."""
    b = require_cuda_f32(b, 'b')
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError('cg: x must be a contiguous float32 CUDA tensor (updated in place)')
    if tuple(b.shape) != A.dim_y or tuple(x.shape) != A.dim_y:
        raise ValueError('cg: b/x shape does not match the operator')
    if max_iter > _lib.UR_CG_MAX_ITER:
        raise NotImplementedError('max_iter > %d' % _lib.UR_CG_MAX_ITER)
    if ws is None:
        ws = _lib.workspace(A.cg_bytes, b.device, 'cg')
    opts = _lib.ur_cg_opts(int(max_iter), int(rule), float(tolerance or 0.0), int(variant))
    st = stream()
    check(lib.ur_cg_solve(C.byref(A.c), ptr(b), ptr(x), ptr(ws), ws.numel(), C.byref(opts), st))
    # the workspace is shared by consecutive solves: keep a stream-ordered snapshot of the
    # device-side CG state (head of the workspace) for this solve's result handle
    snap = ws[:_STATE_BYTES].clone()
    return CgInfo(snap, st)


def cg(A, b, x=None, precond=None, max_iter=None, tolerance=1e-5, verbose=False,
       sum_dtype=torch.float64, inplace=True, stop='E'):
    """Solve A x = b by conjugate gradients (x updated in place when given)."""
    if sum_dtype != torch.float64:
        raise NotImplementedError('dot products are accumulated in float64')
    max_iter = max_iter or len(b) * 10
    if x is None:
        x = torch.zeros_like(b)
    elif not inplace:
        x = x.clone()
    rule = stop_rule(stop, tolerance, verbose)

    if isinstance(A, LhsOperator) and precond is None and not verbose:
        cg.last = cg_fused(A, b, x, max_iter, tolerance, rule)
        return x

    # ---- generic host-driven loop over the CUDA vector kernels ----
    if isinstance(A, torch.Tensor):
        mat = A
        A = lambda v: mat.mm(v)
    b = require_cuda_f32(b, 'b')
    n = b.numel()
    st = stream
    scal = torch.zeros(4, dtype=torch.float64, device=b.device)  # rz, pAp, alpha, beta
    r = b - A(x)
    z = r if precond is None else precond(r)
    check(lib.ur_dot(ptr(r), ptr(z), n, ptr(scal[0:1]), st()))
    p = z.clone()

    def objective():
        if rule == _lib.UR_STOP_RESIDUAL:
            return torch.sqrt(scal[0]).clone()
        e = torch.zeros(1, dtype=torch.float64, device=b.device)
        t = A(x).sub_(2 * b)
        check(lib.ur_dot(ptr(t), ptr(x), n, ptr(e), st()))
        return 0.5 * e[0]

    track = rule != _lib.UR_STOP_NONE
    if track:
        obj = torch.zeros(max_iter + 1, dtype=torch.float64, device=b.device)
        obj[0] = objective()
    n_done = 0
    for it in range(1, max_iter + 1):
        Ap = require_cuda_f32(A(p), 'A(p)')
        check(lib.ur_dot(ptr(p), ptr(Ap), n, ptr(scal[1:2]), st()))
        scal[2] = scal[0] / scal[1]
        rz0 = scal[0].clone()
        if precond is None:
            check(lib.ur_cg_update_xr(ptr(x), ptr(r), ptr(p), ptr(Ap), n, ptr(scal[2:3]),
                                      ptr(scal[0:1]), st()))
            z = r
        else:
            dummy = torch.zeros(1, dtype=torch.float64, device=b.device)
            check(lib.ur_cg_update_xr(ptr(x), ptr(r), ptr(p), ptr(Ap), n, ptr(scal[2:3]),
                                      ptr(dummy), st()))
            z = precond(r)
            check(lib.ur_dot(ptr(r), ptr(z), n, ptr(scal[0:1]), st()))
        scal[3] = scal[0] / rz0
        check(lib.ur_cg_update_p(ptr(p), ptr(z), n, ptr(scal[3:4]), st()))
        n_done = it
        if track:
            obj[it] = objective()
            gain = get_gain(obj[:it + 1], monotonicity='decreasing')
            if verbose:
                print('{:3d} | obj = {:12.6g} | gain = {:12.6g}'.format(
                    it, obj[it].item(), gain.item()))
            if gain.abs() < tolerance:
                break
    cg.last = type('CgHostInfo', (), {'n_iter': n_done,
                                      'obj': obj[:n_done + 1].tolist() if track else None})()
    return x


cg.last = None
