"""The C-ABI library loads without a GPU and exports every function include/b200/b200.h declares. No compute calls."""
import ctypes
import os
import re

import alpaka_b200
from alpaka_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "b200", "b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text))
    typedefs = set(re.findall(r"typedef[^;]*\(\*\s*(b200_[a-z0-9_]+)\s*\)", text))
    return sorted(names - typedefs)


def test_every_declared_entry_point_is_exported():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 70, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in b200.h but not exported: {missing}"


def test_abi_version_and_error_names_without_a_device():
    lib = _lib.load()
    lib.b200_abi_version.restype = ctypes.c_int
    assert lib.b200_abi_version() == 1
    lib.b200_error_name.restype = ctypes.c_char_p
    lib.b200_error_name.argtypes = [ctypes.c_int]
    assert lib.b200_error_name(-1) == b"B200_EINVAL"
    assert lib.b200_error_name(-2) == b"B200_EALIGN"
    assert lib.b200_error_name(0) is not None


def test_no_cpu_fallback_device_calls_fail_loudly_without_gpu():
    """On a box without a device the runtime entries fail with a CUDA error code (never silently succeed)."""
    import torch

    if torch.cuda.is_available():
        return
    lib = _lib.load()
    n = ctypes.c_int(-1)
    rc = lib.b200_device_count(ctypes.byref(n))
    assert rc != 0 or n.value == 0
    s = ctypes.c_void_p()
    assert lib.b200_stream_create(0, ctypes.byref(s)) != 0
    lib.b200_last_error_string.restype = ctypes.c_char_p
    assert b"returned error" in lib.b200_last_error_string() or b"failed" in lib.b200_last_error_string()


def test_package_never_imports_the_oracle():
    """Product code must not reach into oracle/ (checker only)."""
    pkg = os.path.dirname(alpaka_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "libalpaka_ref" not in text, f
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            assert "hotpath_oracle" not in text and "liboracle" not in text, f
