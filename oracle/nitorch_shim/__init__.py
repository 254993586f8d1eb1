"""Pure-PyTorch stand-in for the subset of ``nitorch`` that UniRes' hot path
imports.  TEST INFRASTRUCTURE -- see oracle/__init__.py (parity unpinned)."""
from . import spatial, core, io, plot, tools  # noqa: F401
