"""nitorch.tools._preproc_utils: imported by unires/_core.py:9-19 (co-registration, atlas crop, mean
space -- outside the ADMM/CG hot path); import-only."""


def _stub(name):
    def f(*args, **kwargs):
        raise NotImplementedError('nitorch.tools.%s is outside the ADMM/CG hot path' % name)
    f.__name__ = name
    return f


_mean_space = _stub('_mean_space')
