"""nitorch.spatial names used by UniRes' hot path."""
from ..spatial import (affine_grid, grid_pull, grid_push, identity_grid,  # noqa: F401
                       voxel_size, im_gradient, im_divergence, AffineGrid)
