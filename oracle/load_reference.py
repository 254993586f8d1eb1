"""Import the reference's OWN control-flow files by path on top of the shim.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The files are looked for under
/root/reference (build container) and then under <repo>/baseline/_ref, the
git-ignored `pip install --no-deps --target baseline/_ref` copy of the
reference that travels to the GPU box with the gpurun snapshot (recorded in
DESIGN.md); tests skip when ``available()`` is False.  No reference source is
copied into the repository's history: the modules are executed in place.

    ref = load_reference()          # namespace with .struct ._project ._update
    ref._project._proj_apply('A', dat, po)
"""
import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    cands = [os.environ.get('UNIRES_REFERENCE'), '/root/reference',
             os.path.join(_REPO, 'baseline', '_ref')]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, 'unires', '_project.py')):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'unires', '_project.py'))


def load_by_path(pkg_name, mods=('struct', '_util', '_project', '_update', '_core', 'run')):
    """Execute the reference's files in place as package `pkg_name` on top of whatever
    `nitorch` is registered in sys.modules at this moment (the oracle shim, or the
    product's nitorch_compat for the drop-in tests)."""
    if not available():
        raise FileNotFoundError('reference not found under ' + REFERENCE_ROOT)
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'unires')]
    sys.modules[pkg_name] = pkg
    ns = types.SimpleNamespace()
    for mod in mods:
        full = pkg_name + '.' + mod
        spec = importlib.util.spec_from_file_location(
            full, os.path.join(REFERENCE_ROOT, 'unires', mod + '.py'))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        setattr(pkg, mod, m)
        setattr(ns, mod, m)
    return ns


def install_shim():
    """Expose oracle.nitorch_shim as top-level ``nitorch`` in sys.modules."""
    from oracle import nitorch_shim as shim
    names = {
        'nitorch': shim,
        'nitorch.spatial': shim.spatial,
        'nitorch.core': shim.core,
        'nitorch.core.kernels': shim.core.kernels,
        'nitorch.core.optim': shim.core.optim,
        'nitorch.core.math': shim.core.math,
        'nitorch.core._linalg_expm': shim.core._linalg_expm,
        'nitorch.core.constants': shim.core.constants,
        'nitorch.core.utils': shim.core.utils,
        'nitorch.tools': shim.tools,
        'nitorch.tools.preproc': shim.tools.preproc,
        'nitorch.tools.img_statistics': shim.tools.img_statistics,
        'nitorch.tools._preproc_fov': shim.tools._preproc_fov,
        'nitorch.tools._preproc_utils': shim.tools._preproc_utils,
        'nitorch.io': shim.io,
        'nitorch.plot': shim.plot,
        'nitorch.plot.volumes': shim.plot.volumes,
    }
    for k, v in names.items():
        sys.modules.setdefault(k, v)
    return shim


_cache = None


def load_reference():
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError('reference not found under ' + REFERENCE_ROOT)
    install_shim()
    _cache = load_by_path('_unires_reference')
    return _cache
