"""nitorch.tools.preproc: imported by unires/_core.py:9-19 (co-registration, atlas crop, mean
space -- outside the ADMM/CG hot path); import-only."""


def _stub(name):
    def f(*args, **kwargs):
        raise NotImplementedError('nitorch.tools.%s is outside the ADMM/CG hot path' % name)
    f.__name__ = name
    return f


atlas_crop = _stub('atlas_crop')
affine_align = _stub('affine_align')
atlas_align = _stub('atlas_align')
reset_origin = _stub('reset_origin')
