"""CPU: the C-ABI library builds, loads and exports every symbol include/mridc_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "mridc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from mridc_b200 import _lib

    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header but not exported" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert _lib.load().mrb_version() >= 100


def test_error_codes_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from mridc_b200 import _lib

    lib = _lib.load()
    rc = lib.mrb_fft2_c2c(None, None, 1, 4, 4, 0, 0, 0, None)
    assert rc == -1 and b"null" in lib.mrb_last_error()
    rc = lib.mrb_conv2d(ctypes.c_void_p(8), 0, ctypes.c_void_p(8), None, ctypes.c_void_p(8), 0, 1, 1, 1, 4, 4, 2, 1,
                        0, 0, 0.0, None, None, None, 0, None)
    assert rc == -1 and b"odd" in lib.mrb_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert lib.mrb_dc_workspace_bytes(1, 15, 320, 320) == 2 * 15 * 320 * 320 * 8


def test_missing_library_fails_loudly(monkeypatch):
    from mridc_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmridc_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
