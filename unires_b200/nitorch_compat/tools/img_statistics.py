"""nitorch.tools.img_statistics names used by UniRes (unires/_core.py:15, 122)."""
from ...stats import estimate_noise  # noqa: F401


def estimate_fwhm(*args, **kwargs):
    raise NotImplementedError('estimate_fwhm is outside the ADMM/CG hot path')
