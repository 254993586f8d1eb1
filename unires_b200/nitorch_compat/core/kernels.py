"""nitorch.core.kernels names used by UniRes' hot path."""
from ...kernels import smooth  # noqa: F401
