"""Import the reference's OWN control-flow files by path on top of the shim.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable in the build
container, where /root/reference exists; the GPU box does not have it, so
nothing that runs there may call this (tests skip when ``available()`` is
False).  No reference source is copied: the modules are executed from
/root/reference/unires/*.py in place.

    ref = load_reference()          # namespace with .struct ._project ._update
    ref._project._proj_apply('A', dat, po)
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('UNIRES_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'unires', '_project.py'))


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
    pkg_name = '_unires_reference'
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'unires')]
    sys.modules[pkg_name] = pkg
    ns = types.SimpleNamespace()
    for mod in ('struct', '_util', '_project', '_update', '_core', 'run'):
        full = pkg_name + '.' + mod
        spec = importlib.util.spec_from_file_location(
            full, os.path.join(REFERENCE_ROOT, 'unires', mod + '.py'))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        setattr(pkg, mod, m)
        setattr(ns, mod, m)
    _cache = ns
    return ns
