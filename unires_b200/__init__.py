"""unires_b200 -- B200-native (sm_100a) implementation of UniRes' ADMM /
conjugate-gradient hot path behind the reference's own operator surface.

Layout (only what the path needs):
    csrc/            CUDA kernels + the C ABI (include/unires_b200.h)
    _lib.py          ctypes binding of libunires_b200.so (fails loudly if missing)
    spatial.py       nitorch.spatial drop-ins (affine_grid, grid_pull/push, im_*)
    kernels.py       nitorch.core.kernels.smooth drop-in
    optim.py         nitorch.core.optim drop-ins (cg, get_gain)
    nitorch_compat/  the above arranged under nitorch's module names
    struct.py        unires.struct field-compatible containers
    _project.py      unires._project mirror (_proj, _proj_apply, _proj_info, _DtD, ...)
    _update.py       unires._update mirror (_update_admm, _compute_nll, ...)
    synth.py         synthetic workloads for bench/tests

Importing the package does not touch the GPU; the CUDA library is loaded on
first import of any operator module.
"""
__version__ = '0.1.0'
