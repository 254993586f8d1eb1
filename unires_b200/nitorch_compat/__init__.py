"""The hot-path operators under nitorch's module names, so that UniRes'
imports (unires/_project.py:1-3, unires/_update.py:5-11, unires/_util.py:2-4,
unires/run.py:6-9, unires/_core.py:7-19) resolve to the sm_100a kernels:

    import unires_b200.nitorch_compat as nc
    nc.install()          # registers nitorch, nitorch.spatial, nitorch.core.*, ...
    import unires         # the reference's own files now run on unires_b200

Everything on the ADMM/CG path is backed by a kernel of this package; the
pre-processing helpers UniRes imports at module load (co-registration, atlas crop,
mean space: SURVEY.md section 2 #9, out of scope) are import-only stubs that raise
NotImplementedError when called.  tests/test_compat_dropin.py imports the reference's
unmodified files on top of this and runs its `_update_admm` against the product's.
"""
import sys

from . import spatial, core, tools, plot, io  # noqa: F401

_MODULES = {
    'nitorch.spatial': spatial,
    'nitorch.core': core,
    'nitorch.core.kernels': core.kernels,
    'nitorch.core.optim': core.optim,
    'nitorch.core.math': core.math,
    'nitorch.core._linalg_expm': core._linalg_expm,
    'nitorch.core.constants': core.constants,
    'nitorch.core.utils': core.utils,
    'nitorch.tools': tools,
    'nitorch.tools.img_statistics': tools.img_statistics,
    'nitorch.tools.preproc': tools.preproc,
    'nitorch.tools._preproc_fov': tools._preproc_fov,
    'nitorch.tools._preproc_utils': tools._preproc_utils,
    'nitorch.plot': plot,
    'nitorch.plot.volumes': plot.volumes,
    'nitorch.io': io,
}


def install(force=False):
    """Register this package as `nitorch` in sys.modules (no-op if a real
    nitorch is already imported, unless force=True)."""
    if 'nitorch' in sys.modules and not force:
        return False
    sys.modules['nitorch'] = sys.modules[__name__]
    for name, mod in _MODULES.items():
        sys.modules[name] = mod
    return True


def uninstall():
    """Remove the registrations made by install() (tests)."""
    me = sys.modules[__name__]
    if sys.modules.get('nitorch') is me:
        del sys.modules['nitorch']
    for name, mod in _MODULES.items():
        if sys.modules.get(name) is mod:
            del sys.modules[name]
