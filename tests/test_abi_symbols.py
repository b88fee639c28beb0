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


def test_argument_validation_of_the_fused_entries_needs_no_device():
    """Bad arguments are refused (B200_EINVAL = -1) before any CUDA call: null plans, missing exchange descriptors, slab
    plans that are not slabs, too many ranks. Also true on a box without a GPU."""
    import numpy as np

    lib = _lib.load()
    EINVAL = -1
    tf = (ctypes.c_double * 4)(1.0, 1.0, 1.0, 1.0)
    assert lib.b200_heat2d_step2_f64(None, None, 0, 0.1, 0.1, 1.0, 1.0) == EINVAL
    assert lib.b200_heat2d_stepn_f64(None, None, 0, 0.1, 0.1, 3, tf) == EINVAL
    assert lib.b200_heat2d_step2_halo_f64(None, None, 0, 0.1, 0.1, 1.0, 1.0, 1) == EINVAL
    assert lib.b200_heat2d_stepn_halo_f64(None, None, 0, 0.1, 0.1, 4, tf, 1) == EINVAL
    # a slab keeps the full width (LEFT | RIGHT physical), at least 2 * ghost rows, ghost depth 2..8
    sx, sy = np.zeros(66), np.zeros(72)
    plan = ctypes.c_void_p()
    fake = ctypes.c_void_p(256)  # never dereferenced: the checks below fail first
    for edges, ny, ghost in ((1 | 2, 64, 2), (4 | 8, 3, 2), (4 | 8, 64, 1), (4 | 8, 64, 9)):
        rc = lib.b200_heat2d_slab_plan_create(0, fake, fake, 66 * 8 + 16, ny, 64, sx.ctypes.data, sy.ctypes.data, edges, ghost,
                                              ctypes.byref(plan))
        assert rc == EINVAL, (edges, ny, ghost, rc)
    out = ctypes.c_void_p(256)
    assert lib.b200_dot_allranks_f64(None, fake, fake, 16, out, fake, None, 1) == EINVAL  # no exchange descriptor
    ex = _lib.Exchange()
    ex.world, ex.rank = 17, 0  # more ranks than B200_EXCHANGE_MAX_RANKS
    assert lib.b200_reduce_sum_allranks_u32(None, fake, 16, out, fake, ctypes.byref(ex), 1) == EINVAL
    ex.world, ex.rank = 2, 0  # step 0 is not a call number; base[1] missing
    assert lib.b200_reduce_sum_allranks_u32(None, fake, 16, out, fake, ctypes.byref(ex), 0) == EINVAL
    ex.base[0] = 256
    assert lib.b200_reduce_sum_allranks_u32(None, fake, 16, out, fake, ctypes.byref(ex), 1) == EINVAL
