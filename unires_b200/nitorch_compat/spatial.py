"""nitorch.spatial names used by UniRes (unires/_project.py:2-3, unires/_update.py:5-7,
unires/run.py:6, unires/_core.py:7-8)."""
import torch

from ..spatial import (affine_grid, grid_pull, grid_push, identity_grid,  # noqa: F401
                       voxel_size, im_gradient, im_divergence, AffineGrid, affine_grad)


def grid_grad(input, grid, interpolation='linear', bound='zero', extrapolate=False):
    """nitorch.spatial.grid_grad as called at unires/_update.py:505: spatial gradient of the
    trilinearly interpolated (1, 1, X, Y, Z) `input` at the grid points, (1, 1, *grid, 3).
    UniRes only ever passes the affine grid it has just built (unires/_update.py:498-499), so
    the coordinates are evaluated in-kernel from the 3x4 matrix (`ur_affine_grad`); a dense
    coordinate field has no derivative kernel here."""
    if interpolation not in ('linear', 1):
        raise NotImplementedError('grid_grad: interpolation %r' % (interpolation,))
    if bound not in ('zero', 'zeros'):
        raise NotImplementedError('grid_grad: bound %r' % (bound,))
    if not isinstance(grid, AffineGrid):
        raise NotImplementedError('grid_grad: only lazy affine grids (affine_grid(...)) are '
                                  'supported')
    if input.dim() != 5 or input.shape[0] != 1 or input.shape[1] != 1:
        raise ValueError('grid_grad: input must be (1, 1, X, Y, Z)')
    mat = torch.tensor(grid.rows, dtype=torch.float32).reshape(3, 4)
    g = affine_grad(input[0, 0], mat, grid.shape_, extrapolate=extrapolate)
    return g[None, None]


def _out_of_scope(name):
    def stub(*args, **kwargs):
        raise NotImplementedError('nitorch.spatial.%s is outside the ADMM/CG hot path '
                                  '(SURVEY.md section 2, #9)' % name)
    stub.__name__ = name
    return stub


# imported by unires/_core.py:7-8 (co-registration / mean space: out of scope, import-only)
affine_matrix_classic = _out_of_scope('affine_matrix_classic')
affine_basis = _out_of_scope('affine_basis')
max_bb = _out_of_scope('max_bb')
