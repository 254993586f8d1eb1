from . import kernels, optim  # noqa: F401
