from . import img_statistics, preproc, _preproc_fov, _preproc_utils  # noqa: F401
