"""Pure-PyTorch restatement of the nitorch.spatial functions UniRes calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: restates the
published nitorch algorithms; anchored on the
reference call sites:
  affine_grid    unires/_project.py:159
  grid_pull      unires/_project.py:164,174,183,187
  grid_push      unires/_project.py:172,179,185,188
  im_gradient    unires/_project.py:314, unires/_update.py:168,176,188,419
  im_divergence  unires/_project.py:315, unires/_update.py:132
  voxel_size     unires/_project.py:224,230, unires/_update.py:111
"""
import torch

# in-FOV tolerance used by nitorch when extrapolate=False (Appendix A.2, Q2)
FOV_TOLERANCE = 5e-2


def voxel_size(mat):
    """Column norms of the 3x3 part of an affine matrix (Appendix A, 8b)."""
    dim = mat.shape[-1] - 1
    return (mat[:dim, :dim] ** 2).sum(0).sqrt()


def identity_grid(shape, dtype=None, device=None, jitter=False):
    """Voxel-coordinate identity grid, shape (*shape, 3)."""
    if jitter:
        raise NotImplementedError('jitter')
    mesh = [torch.arange(s, dtype=dtype, device=device) for s in shape]
    return torch.stack(torch.meshgrid(*mesh, indexing='ij'), dim=-1)


def affine_grid(mat, shape, jitter=False):
    """grid[i,j,k,:] = mat[:3,:3] @ (i,j,k) + mat[:3,3], in mat.dtype (A.1)."""
    nb_dim = mat.shape[-1] - 1
    grid = identity_grid(shape, dtype=mat.dtype, device=mat.device, jitter=jitter)
    lin = mat[:nb_dim, :nb_dim]
    off = mat[:nb_dim, -1]
    # same association as a matvec per voxel: sum_j lin[i, j] * g[j] + off[i]
    grid = torch.matmul(grid, lin.transpose(0, 1)) + off
    return grid


def _check_opts(bound, interpolation):
    if bound not in ('zero', 'zeros'):
        raise NotImplementedError('oracle shim only restates bound="zero"')
    if interpolation in ('linear', 1):
        return 1
    if interpolation in ('nearest', 0):
        return 0
    raise NotImplementedError('oracle shim only restates interpolation 0/1')


def _inbounds(g, shape, extrapolate):
    if extrapolate:
        return None
    t = FOV_TOLERANCE
    msk = torch.ones(g.shape[:-1], dtype=torch.bool, device=g.device)
    for d, n in enumerate(shape):
        msk &= (g[..., d] > -t) & (g[..., d] < n - 1 + t)
    return msk


def _corners(g, shape, order):
    """Yield (flat_index, weight, valid) for every interpolation corner."""
    nd = len(shape)
    if order == 0:
        idx = torch.floor(g + 0.5).long()
        valid = torch.ones(g.shape[:-1], dtype=torch.bool, device=g.device)
        flat = torch.zeros(g.shape[:-1], dtype=torch.long, device=g.device)
        for d, n in enumerate(shape):
            i = idx[..., d]
            valid &= (i >= 0) & (i < n)
            flat = flat * n + i.clamp(0, n - 1)
        yield flat, None, valid
        return
    g0 = torch.floor(g)
    w1 = g - g0
    w0 = 1 - w1
    i0 = g0.long()
    for corner in range(2 ** nd):
        flat = torch.zeros(g.shape[:-1], dtype=torch.long, device=g.device)
        wgt = None
        valid = torch.ones(g.shape[:-1], dtype=torch.bool, device=g.device)
        for d, n in enumerate(shape):
            bit = (corner >> (nd - 1 - d)) & 1
            i = i0[..., d] + bit
            w = w1[..., d] if bit else w0[..., d]
            valid &= (i >= 0) & (i < n)
            flat = flat * n + i.clamp(0, n - 1)
            wgt = w if wgt is None else wgt * w
        yield flat, wgt, valid


def grid_pull(input, grid, interpolation='linear', bound='zero', extrapolate=False):
    """Trilinear gather (A.2).  input (B,C,*in), grid (B,*out,3) -> (B,C,*out)."""
    order = _check_opts(bound, interpolation)
    B, C = input.shape[:2]
    shape = tuple(input.shape[2:])
    oshape = tuple(grid.shape[1:-1])
    out = input.new_zeros((B, C) + oshape)
    for b in range(B):
        g = grid[b].reshape(-1, len(shape)).to(input.dtype)
        msk = _inbounds(g, shape, extrapolate)
        src = input[b].reshape(C, -1)
        acc = input.new_zeros((C, g.shape[0]))
        for flat, wgt, valid in _corners(g, shape, order):
            if msk is not None:
                valid = valid & msk
            v = src[:, flat]
            if wgt is not None:
                v = v * wgt
            acc += v * valid.to(input.dtype)
        out[b] = acc.reshape((C,) + oshape)
    return out


def grid_push(input, grid, shape=None, interpolation='linear', bound='zero',
              extrapolate=False):
    """Exact transpose of grid_pull (A.3).  input (B,C,*out) -> (B,C,*shape)."""
    order = _check_opts(bound, interpolation)
    B, C = input.shape[:2]
    if shape is None:
        shape = tuple(input.shape[2:])
    shape = tuple(int(s) for s in shape)
    n_tgt = 1
    for s in shape:
        n_tgt *= s
    out = input.new_zeros((B, C, n_tgt))
    for b in range(B):
        g = grid[b].reshape(-1, len(shape)).to(input.dtype)
        msk = _inbounds(g, shape, extrapolate)
        src = input[b].reshape(C, -1)
        for flat, wgt, valid in _corners(g, shape, order):
            if msk is not None:
                valid = valid & msk
            v = src
            if wgt is not None:
                v = v * wgt
            v = v * valid.to(input.dtype)
            out[b].index_add_(1, flat, v)
    return out.reshape((B, C) + shape)


def grid_grad(input, grid, interpolation='linear', bound='zero', extrapolate=False):
    """Spatial gradient of the trilinearly interpolated volume with respect to the voxel
    coordinates, sampled at `grid` (unires/_update.py:505).  input (B,C,*in), grid (B,*out,3)
    -> (B,C,*out,3).  Along axis d the two corner weights (1 - t, t) become (-1, +1); corners
    outside the volume contribute zero and samples outside the field of view are zero, like
    grid_pull (A.2).  [EXT-UNVERIFIED] restated from nitorch's published behaviour."""
    order = _check_opts(bound, interpolation)
    if order != 1:
        raise NotImplementedError('grid_grad: linear interpolation only')
    B, C = input.shape[:2]
    shape = tuple(input.shape[2:])
    nd = len(shape)
    oshape = tuple(grid.shape[1:-1])
    out = input.new_zeros((B, C) + oshape + (nd,))
    for b in range(B):
        g = grid[b].reshape(-1, nd).to(input.dtype)
        msk = _inbounds(g, shape, extrapolate)
        src = input[b].reshape(C, -1)
        g0 = torch.floor(g)
        w1 = g - g0
        w0 = 1 - w1
        i0 = g0.long()
        acc = input.new_zeros((C, g.shape[0], nd))
        for corner in range(2 ** nd):
            flat = torch.zeros(g.shape[0], dtype=torch.long, device=g.device)
            valid = torch.ones(g.shape[0], dtype=torch.bool, device=g.device)
            bits = [(corner >> (nd - 1 - d)) & 1 for d in range(nd)]
            for d, n in enumerate(shape):
                i = i0[:, d] + bits[d]
                valid &= (i >= 0) & (i < n)
                flat = flat * n + i.clamp(0, n - 1)
            if msk is not None:
                valid = valid & msk
            v = src[:, flat] * valid.to(input.dtype)
            for a in range(nd):  # derivative along axis a
                wgt = None
                for d in range(nd):
                    if d == a:
                        w = torch.full_like(w1[:, d], 1.0 if bits[d] else -1.0)
                    else:
                        w = w1[:, d] if bits[d] else w0[:, d]
                    wgt = w if wgt is None else wgt * w
                acc[:, :, a] += v * wgt
        out[b] = acc.reshape((C,) + oshape + (nd,))
    return out


def _vx(vx, ref):
    vx = torch.as_tensor(vx, dtype=ref.dtype, device=ref.device).flatten()
    if vx.numel() == 1:
        vx = vx.expand(3)
    return vx


def im_gradient(dat, vx=None, which='forward', bound='zero'):
    """Forward finite differences, zero past the high edge (A.4)."""
    if which != 'forward' or bound not in ('zero', 'zeros'):
        raise NotImplementedError
    vx = _vx(1.0 if vx is None else vx, dat)
    nd = 3
    out = []
    for a in range(nd):
        ax = dat.dim() - nd + a
        pad = [0, 0] * nd
        pad[2 * (nd - 1 - a) + 1] = 1  # high side of axis a
        p = torch.nn.functional.pad(dat, pad)
        hi = p.narrow(ax, 1, dat.shape[ax])
        out.append((hi - dat) / vx[a])
    return torch.stack(out, dim=0)


def im_divergence(dat, vx=None, which='forward', bound='zero'):
    """Transpose of im_gradient: zero before the low edge (A.5)."""
    if which != 'forward' or bound not in ('zero', 'zeros'):
        raise NotImplementedError
    vx = _vx(1.0 if vx is None else vx, dat)
    nd = 3
    out = None
    for a in range(nd):
        d = dat[a]
        ax = d.dim() - nd + a
        pad = [0, 0] * nd
        pad[2 * (nd - 1 - a)] = 1  # low side of axis a
        p = torch.nn.functional.pad(d, pad)
        n = d.shape[ax]
        t = (p.narrow(ax, 0, n) - p.narrow(ax, 1, n)) / vx[a]
        out = t if out is None else out + t
    return out


def affine_basis(group='SE', dim=3, dtype=None, device=None):
    """Lie-algebra basis of SE(3): 3 translations then 3 rotations (unires/_core.py:317)."""
    if group != 'SE' or dim != 3:
        raise NotImplementedError
    B = torch.zeros(6, 4, 4, dtype=dtype or torch.float64, device=device)
    for i in range(3):
        B[i, i, 3] = 1
    for k, (i, j) in enumerate(((0, 1), (0, 2), (1, 2))):
        B[3 + k, i, j] = 1
        B[3 + k, j, i] = -1
    return B


def affine_matrix_classic(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('affine_matrix_classic is only used by pre-processing')


def max_bb(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('max_bb is only used by pre-processing')
