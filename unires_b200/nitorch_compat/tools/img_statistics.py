"""nitorch.tools.img_statistics names used by UniRes (unires/_core.py:15, 122)."""
from ...stats import estimate_noise  # noqa: F401
