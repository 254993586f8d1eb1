"""Import-only stub (pre-processing is out of scope; unires/_core.py imports these names)."""


def _unavailable(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError('nitorch.tools is out of scope (SURVEY.md section 2)')


atlas_crop = affine_align = atlas_align = reset_origin = _unavailable
estimate_fwhm = estimate_noise = _unavailable
_bb_atlas = _mean_space = _unavailable
