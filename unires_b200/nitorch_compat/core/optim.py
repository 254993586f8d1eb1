"""nitorch.core.optim names used by UniRes' hot path."""
from ...optim import cg, get_gain  # noqa: F401
