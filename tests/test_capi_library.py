"""The C-ABI library: loads, exports every symbol include/unires_b200.h declares, and the
ctypes structure layouts agree with the C compiler's.  No compute calls (no GPU needed)."""
import ctypes
import os
import subprocess
import tempfile

from unires_b200 import _lib


def test_library_exports_every_declared_symbol():
    declared = _lib.header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(_lib.lib, name), 'missing export: ' + name
    assert set(declared) == set(_lib._SIGNATURES), 'ctypes signature table out of date'
    assert _lib.lib.ur_version() >= 100


def test_struct_layouts_match_c():
    src = '#include <stdio.h>\n#include "unires_b200.h"\nint main(){printf("%zu %zu %zu\\n",' \
          'sizeof(ur_proj),sizeof(ur_lhs),sizeof(ur_cg_opts));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 't.c')
        with open(c, 'w') as f:
            f.write(src)
        exe = os.path.join(d, 't')
        subprocess.check_call(['gcc', '-I', os.path.dirname(_lib.HEADER_PATH), c, '-o', exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.ur_proj), ctypes.sizeof(_lib.ur_lhs),
                     ctypes.sizeof(_lib.ur_cg_opts)]


def test_error_codes_map_to_reference_exceptions():
    import pytest
    # a NULL operator is an argument error -> ValueError, like unires/_project.py:123-126
    rc = _lib.lib.ur_proj_apply(7, None, None, None, None, 0, None)
    assert rc == _lib.UR_ERR_ARG
    with pytest.raises(ValueError, match='Undefined operator'):
        _lib.check(rc)
