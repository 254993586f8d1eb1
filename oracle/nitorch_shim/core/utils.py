"""nitorch.core.utils.ceil_pow stub (unires/_core.py:17; only used by the pow-crop)."""


def ceil_pow(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('ceil_pow is out of scope')
