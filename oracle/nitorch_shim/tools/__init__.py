"""Import-only stubs for nitorch.tools (pre-processing: out of scope).  TEST INFRASTRUCTURE."""
from . import preproc, img_statistics, _preproc_fov, _preproc_utils  # noqa: F401
