from . import img_statistics  # noqa: F401
