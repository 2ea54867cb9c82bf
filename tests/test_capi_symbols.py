"""CPU: the C-ABI library builds, loads and exports every symbol include/ltb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from livetalking_b200 import build
    return build.build()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ltb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(built_lib):
    from livetalking_b200 import _capi
    names = _header_functions()
    assert len(names) >= 20
    assert sorted(_capi.EXPORTED_SYMBOLS) == names
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), n


def test_library_is_sm100a_tensor_core_code(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass          # tcgen05.mma
    assert "LDTM" in sass             # tcgen05.ld


def test_version_and_error_string(built_lib):
    from livetalking_b200 import _capi
    lib = _capi.lib()
    assert lib.ltb_version() >= 100
    # a failing call must set a message and must not abort: null session
    assert lib.ltb_w2l_sync(None) != 0
    assert b"null" in lib.ltb_last_error()
