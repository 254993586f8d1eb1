"""nitorch.spatial drop-ins backed by the sm_100a kernels (no CPU fallback).

Same call signatures as the functions UniRes imports at
unires/_project.py:2-3 and unires/_update.py:5-7:

    affine_grid, grid_pull, grid_push, identity_grid, voxel_size,
    im_gradient, im_divergence

`affine_grid` returns a lazy :class:`AffineGrid` (3x4 matrix + shape).  When it
is handed to grid_pull / grid_push the coordinates are evaluated inside the
kernel and the dense (X,Y,Z,3) field of unires/_project.py:159 is never
written; indexing it any other way (unires/run.py:169-174) materialises it.
"""
import torch

from . import _lib
from ._lib import lib, check, ptr, i3, f3, stream, require_cuda_f32

_ORDERS = {'linear': 1, 1: 1, 'nearest': 0, 0: 0}


def _order(interpolation):
    try:
        return _ORDERS[interpolation]
    except (KeyError, TypeError):
        raise NotImplementedError('interpolation %r: only orders 0 and 1 are implemented'
                                  % (interpolation,))


def _bound(bound):
    if bound not in ('zero', 'zeros'):
        raise NotImplementedError('bound %r: only "zero" is implemented' % (bound,))


def _which(which):
    if which != 'forward':
        raise NotImplementedError('finite difference %r: only "forward" is implemented'
                                  % (which,))


def voxel_size(mat):
    """Column norms of the linear part of an affine matrix."""
    nd = mat.shape[-1] - 1
    return (mat[:nd, :nd] ** 2).sum(0).sqrt()


def identity_grid(shape, dtype=None, device=None, jitter=False):
    if jitter:
        raise NotImplementedError('jitter')
    axes = [torch.arange(s, dtype=dtype, device=device) for s in shape]
    return torch.stack(torch.meshgrid(*axes, indexing='ij'), dim=-1)


class AffineGrid:
    """Lazy voxel-coordinate grid ``grid[i,j,k] = M[:3,:3] @ (i,j,k) + M[:3,3]``."""

    def __init__(self, mat, shape, batched=False):
        self.mat = mat
        self.shape_ = tuple(int(s) for s in shape)
        self.batched = batched
        m = mat.detach().to('cpu', torch.float32)  # float32 like the reference's cast
        self.rows = [float(v) for v in m[:3, :].reshape(-1)]
        self.dtype = mat.dtype
        self.device = mat.device

    @property
    def shape(self):
        lead = (1,) if self.batched else ()
        return torch.Size(lead + self.shape_ + (3,))

    def materialize(self):
        if not self.mat.is_cuda:
            raise RuntimeError('unires_b200: affine_grid can only be materialised on a CUDA '
                               'device (no CPU fallback)')
        out = torch.empty(self.shape_ + (3,), dtype=torch.float32, device=self.device)
        check(lib.ur_affine_grid(_lib.farr(self.rows), ptr(out), i3(self.shape_), stream()))
        out = out.to(self.dtype)
        return out[None] if self.batched else out

    def __getitem__(self, index):
        if not self.batched and (index is None or index == (None, Ellipsis)
                                 or index == (None,)):
            return AffineGrid(self.mat, self.shape_, batched=True)
        return self.materialize()[index]


def affine_grid(mat, shape, jitter=False):
    if jitter:
        raise NotImplementedError('jitter')
    return AffineGrid(mat, shape)


def _spatial(t, what):
    if t.dim() != 5:
        raise ValueError('%s must be (B, C, X, Y, Z)' % what)
    return tuple(t.shape[2:])


def _resample(input, grid, target_shape, interpolation, bound, extrapolate, push):
    _bound(bound)
    order = _order(interpolation)
    inp = require_cuda_f32(input, 'input')
    B, Cn = inp.shape[:2]
    lazy = isinstance(grid, AffineGrid)
    if lazy:
        gshape = grid.shape_
        if grid.device != inp.device and grid.mat.is_cuda:
            raise RuntimeError('grid and input are on different devices')
    else:
        g = require_cuda_f32(grid.to(torch.float32), 'grid')
        if g.dim() != 5 or g.shape[-1] != 3:
            raise ValueError('grid must be (B, X, Y, Z, 3)')
        gshape = tuple(g.shape[1:4])
    in_sp = _spatial(inp, 'input')
    if push:
        if in_sp != gshape:
            raise ValueError('grid_push: input %s and grid %s spatial shapes differ'
                             % (in_sp, gshape))
        out_sp = tuple(int(s) for s in (target_shape if target_shape is not None else in_sp))
        out = torch.zeros((B, Cn) + out_sp, dtype=torch.float32, device=inp.device)
    else:
        out_sp = gshape
        out = torch.empty((B, Cn) + out_sp, dtype=torch.float32, device=inp.device)
    ext = 1 if extrapolate else 0
    for b in range(B):
        for c in range(Cn):
            src, dst = inp[b, c], out[b, c]
            if lazy:
                mat = _lib.farr(grid.rows)
                if push:
                    check(lib.ur_affine_push(ptr(src), i3(in_sp), mat, ptr(dst), i3(out_sp),
                                             order, ext, 1.0, stream()))
                else:
                    check(lib.ur_affine_pull(ptr(src), i3(in_sp), mat, ptr(dst), i3(out_sp),
                                             order, ext, stream()))
            else:
                gb = g[b if g.shape[0] > 1 else 0]
                if push:
                    check(lib.ur_grid_push(ptr(src), i3(in_sp), ptr(gb), ptr(dst), i3(out_sp),
                                           order, ext, 1.0, stream()))
                else:
                    check(lib.ur_grid_pull(ptr(src), i3(in_sp), ptr(gb), ptr(dst), i3(out_sp),
                                           order, ext, stream()))
    return out


def grid_pull(input, grid, interpolation='linear', bound='zero', extrapolate=False):
    """Trilinear (or nearest) gather; out-of-FOV samples are zero."""
    return _resample(input, grid, None, interpolation, bound, extrapolate, push=False)


def grid_push(input, grid, shape=None, interpolation='linear', bound='zero',
              extrapolate=False):
    """Exact transpose of :func:`grid_pull` (scatter-add into `shape`)."""
    return _resample(input, grid, shape, interpolation, bound, extrapolate, push=True)


def _vx3(vx):
    if vx is None:
        return [1.0, 1.0, 1.0]
    v = torch.as_tensor(vx).detach().to('cpu', torch.float32).flatten().tolist()
    if len(v) == 1:
        v = v * 3
    if len(v) != 3:
        raise ValueError('vx must have 1 or 3 elements')
    return v


def im_gradient(dat, vx=None, which='forward', bound='zero'):
    """(X,Y,Z) -> (3,X,Y,Z) forward differences / vx, zero past the high edge."""
    _which(which)
    _bound(bound)
    d = require_cuda_f32(dat, 'dat')
    if d.dim() != 3:
        raise ValueError('im_gradient: dat must be (X, Y, Z)')
    out = torch.empty((3,) + tuple(d.shape), dtype=torch.float32, device=d.device)
    check(lib.ur_im_gradient(ptr(d), ptr(out), i3(d.shape), f3(_vx3(vx)), stream()))
    return out


def im_divergence(dat, vx=None, which='forward', bound='zero'):
    """(3,X,Y,Z) -> (X,Y,Z); the transpose of :func:`im_gradient`."""
    _which(which)
    _bound(bound)
    d = require_cuda_f32(dat, 'dat')
    if d.dim() != 4 or d.shape[0] != 3:
        raise ValueError('im_divergence: dat must be (3, X, Y, Z)')
    out = torch.empty(tuple(d.shape[1:]), dtype=torch.float32, device=d.device)
    check(lib.ur_im_divergence(ptr(d), ptr(out), i3(d.shape[1:]), f3(_vx3(vx)), stream()))
    return out


def affine_grad(dat, mat, shape, extrapolate=False):
    """nitorch.spatial.grid_grad for an affine sampling grid (unires/_update.py:505): the
    gradient of the trilinearly interpolated (X, Y, Z) volume `dat` w.r.t. the sampling
    coordinates at grid[i,j,k] = mat[:3,:3] (i,j,k) + mat[:3,3]; returns (*shape, 3)."""
    d = require_cuda_f32(dat, 'dat')
    m = torch.as_tensor(mat).detach().to('cpu', torch.float32)[:3, :].reshape(-1).tolist()
    out = torch.empty(tuple(shape) + (3,), dtype=torch.float32, device=d.device)
    check(lib.ur_affine_grad(ptr(d), i3(d.shape), _lib.farr(m), ptr(out), i3(shape),
                             1 if extrapolate else 0, stream()))
    return out
