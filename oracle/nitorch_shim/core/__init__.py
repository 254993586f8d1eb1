from . import kernels, optim, math, _linalg_expm  # noqa: F401
