"""The hot-path operators under nitorch's module names, so that UniRes'
imports (unires/_project.py:1-3, unires/_update.py:5-9, unires/run.py:6-7)
resolve to the sm_100a kernels:

    import sys, unires_b200.nitorch_compat as nc
    nc.install()          # registers nitorch, nitorch.spatial, nitorch.core.*
    import unires         # now runs on unires_b200

Only the functions on the ADMM/CG path exist, plus
nitorch.tools.img_statistics.estimate_noise (hyper-parameter estimate).
"""
import sys

from . import spatial, core, tools  # noqa: F401


def install(force=False):
    """Register this package as `nitorch` in sys.modules (no-op if a real
    nitorch is already imported, unless force=True)."""
    if 'nitorch' in sys.modules and not force:
        return False
    me = sys.modules[__name__]
    sys.modules['nitorch'] = me
    sys.modules['nitorch.spatial'] = spatial
    sys.modules['nitorch.core'] = core
    sys.modules['nitorch.core.kernels'] = core.kernels
    sys.modules['nitorch.core.optim'] = core.optim
    sys.modules['nitorch.tools'] = tools
    sys.modules['nitorch.tools.img_statistics'] = tools.img_statistics
    return True
