from . import volumes  # noqa: F401
