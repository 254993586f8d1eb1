from . import kernels, optim, math, _linalg_expm, constants, utils  # noqa: F401
