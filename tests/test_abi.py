"""CPU: the C-ABI library loads and exports exactly the symbols include/mfgp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mfgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:mfgp|cov|choi)_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from mfgp_coverage_b200 import _native
    if not os.path.isfile(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(handle, n), n
    assert sorted(_native.SIGNATURES) == names      # the ctypes table binds every declared entry point, and nothing else


def test_host_only_entry_points():
    from mfgp_coverage_b200 import _native
    lib = _native.lib()
    assert b"sm_100a" in lib.mfgp_version()
    assert lib.mfgp_npad(0) == 64 and lib.mfgp_npad(64) == 64 and lib.mfgp_npad(65) == 128
    assert lib.mfgp_workspace_bytes(4096) >= 4096 * 4096 * 2
    assert lib.cov_workspace_bytes(1 << 20, 64, 64) > 0
    assert ctypes.sizeof(_native.MfgpParams) == 88


def test_missing_library_fails_loudly(monkeypatch):
    from mfgp_coverage_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libmfgp_b200.so")
    with pytest.raises(_native.NativeLibraryMissing):
        _native.lib()


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from mfgp_coverage_b200.gaussian_process import SFGP
    with pytest.raises(RuntimeError):
        SFGP(np.zeros((1, 2)), np.zeros((1, 1)), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mfgp-coverage_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_header_prototypes_match_the_ctypes_table():
    """Every prototype of include/mfgp_b200.h against _native.SIGNATURES: same number of arguments and, argument by argument, the
    same class (pointer / double / 64-bit integer / int) -- a drift between the header and the ctypes table would pass garbage
    to the library without any error."""
    from mfgp_coverage_b200 import _native
    text = open(os.path.join(ROOT, "include", "mfgp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)

    def cls_of_c(arg):
        arg = arg.strip()
        if "*" in arg:
            return "ptr"
        if re.match(r"(const\s+)?double\b", arg):
            return "double"
        if re.match(r"(const\s+)?(int64_t|uint64_t)\b", arg):
            return "i64"
        if re.match(r"(const\s+)?(int|int32_t)\b", arg):
            return "int"
        raise AssertionError("unclassified C argument: " + arg)

    def cls_of_ctypes(t):
        if t is ctypes.c_double:
            return "double"
        if t in (ctypes.c_int64, ctypes.c_uint64):
            return "i64"
        if t in (ctypes.c_int, ctypes.c_int32):
            return "int"
        return "ptr"          # c_void_p, c_char_p, POINTER(struct)

    for name, (restype, argtypes) in _native.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;{]*)\)\s*;" % re.escape(name), text)
        assert m, name
        raw = m.group(1).strip()
        args = [] if raw in ("", "void") else [a for a in raw.split(",")]
        assert len(args) == len(argtypes), (name, len(args), len(argtypes))
        for k, (a, t) in enumerate(zip(args, argtypes)):
            assert cls_of_c(a) == cls_of_ctypes(t), (name, k, a.strip(), t)
