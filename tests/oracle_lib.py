"""ctypes loaders for the two CHECKERS under oracle/ (test infrastructure only; never imported by alpaka_b200/):

  oracle()  -> oracle/_build/liboracle.so    plain-C restatement (built on demand with gcc; always available)
  ref()     -> oracle/_ref/libalpaka_ref.so  the unmodified reference compiled from /root/reference (prebuilt;
               None when neither the .so nor /root/reference is present)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libalpaka_ref.so")

_vp, _u64, _u32, _i, _sz, _f64, _f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t, C.c_double, C.c_float

_oracle = None
_ref = None
_ref_tried = False


def _make(target: str) -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, target], check=True, capture_output=True)


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is not None:
        return _oracle
    src = os.path.join(ORACLE_DIR, "hotpath_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        _make("oracle")
    L = C.CDLL(ORACLE_SO)
    for sfx, ft in (("f64", _f64), ("f32", _f32)):
        getattr(L, f"orc_init_{sfx}").argtypes = [_vp, _vp, _vp, ft, _u64]
        getattr(L, f"orc_copy_{sfx}").argtypes = [_vp, _vp, _u64]
        getattr(L, f"orc_mul_{sfx}").argtypes = [_vp, _vp, ft, _u64]
        getattr(L, f"orc_add_{sfx}").argtypes = [_vp, _vp, _vp, _u64]
        getattr(L, f"orc_triad_{sfx}").argtypes = [_vp, _vp, _vp, ft, _u64]
        getattr(L, f"orc_nstream_{sfx}").argtypes = [_vp, _vp, _vp, ft, _u64]
        for k in ("init", "copy", "mul", "add", "triad", "nstream"):
            getattr(L, f"orc_{k}_{sfx}").restype = None
        getattr(L, f"orc_dot_{sfx}").argtypes = [_vp, _vp, _u64, _u32, _u32, _vp]
        getattr(L, f"orc_dot_{sfx}").restype = ft
    L.orc_reduce_block_count.argtypes = [_u64, _u32, _u32]
    L.orc_reduce_block_count.restype = _u32
    for sfx in ("u32", "i32", "u64", "f32", "f64"):
        getattr(L, f"orc_reduce_{sfx}").argtypes = [_vp, _u64, _u32, _u32, _i, _vp]
    L.orc_heat2d_exact.argtypes = [_f64, _f64, _f64]
    L.orc_heat2d_exact.restype = _f64
    L.orc_heat2d_init.argtypes = [_vp, _u32, _u32, _sz, _f64, _f64]
    L.orc_heat2d_init.restype = None
    L.orc_heat2d_validate.argtypes = [_vp, _u32, _u32, _sz, _f64, _f64, _f64]
    L.orc_heat2d_validate.restype = _f64
    L.orc_heat2d_step.argtypes = [_vp, _vp, _u32, _u32, _sz, _u32, _f64, _f64, _f64]
    L.orc_heat2d_step.restype = None
    L.orc_heat2d_run.argtypes = [_vp, _u32, _u32, _u32, _u32, _f64, _f64, _f64]
    L.orc_heat2d_boundary_tables.argtypes = [_vp, _vp, _u32, _u32, _f64, _f64]
    L.orc_heat2d_boundary_tables.restype = None
    L.orc_heat2d_time_factor.argtypes = [_u32, _f64]
    L.orc_heat2d_time_factor.restype = _f64
    for k in ("uniform_f64", "uniform_f32", "hash_u32", "bernoulli_f32"):
        getattr(L, f"orc_fill_{k}").argtypes = [_vp, _u64, _u64, _u64]
        getattr(L, f"orc_fill_{k}").restype = None
    _oracle = L
    return L


def ref():
    """The compiled reference, or None. Built here when /root/reference exists; on the GPU box only the prebuilt
    .so (which travels with the snapshot) is used."""
    global _ref, _ref_tried
    if _ref is not None or _ref_tried:
        return _ref
    _ref_tried = True
    if not os.path.exists(REF_SO):
        if not os.path.isdir("/root/reference/include/alpaka"):
            return None
        try:
            _make("ref")
        except Exception:
            return None
    try:
        L = C.CDLL(REF_SO)
    except OSError:
        return None
    L.ref_babelstream_run.argtypes = [_i, _i, _i, _vp, _vp, _vp, _f64, _u64]
    L.ref_babelstream_dot.argtypes = [_i, _i, _vp, _vp, _u64, _u32, _vp]
    L.ref_babelstream_dot.restype = _f64
    L.ref_babelstream_time.argtypes = [_i, _i, _u64, _i, _vp, _u32]
    L.ref_omp_max_threads.restype = _i
    L.ref_omp_set_threads.argtypes = [_i]
    for sfx in ("u32", "i32", "u64", "f32", "f64"):
        getattr(L, f"ref_reduce_{sfx}").argtypes = [_i, _vp, _u64, _vp, _vp]
    L.ref_heat2d_run.argtypes = [_i, _vp, _u32, _u32, _u32, _u32, _f64, _f64, _f64, _vp]
    L.ref_heat2d_init.argtypes = [_vp, _u32, _u32, _f64, _f64]
    L.ref_heat2d_init.restype = None
    L.ref_heat2d_validate.argtypes = [_vp, _u32, _u32, _f64, _f64, _f64]
    L.ref_heat2d_validate.restype = _f64
    L.ref_heat2d_exact.argtypes = [_f64, _f64, _f64]
    L.ref_heat2d_exact.restype = _f64
    _ref = L
    return L


# ---- small numpy-facing helpers shared by the tests -------------------------------------------------------------
def P(x: np.ndarray):
    return x.ctypes.data_as(C.c_void_p)


SEED = 0x5EED


def fill(kind: str, n: int, seed: int = SEED, first: int = 0) -> np.ndarray:
    dt = {"uniform_f64": np.float64, "uniform_f32": np.float32, "hash_u32": np.uint32, "bernoulli_f32": np.float32}[kind]
    x = np.empty(n, dtype=dt)
    getattr(oracle(), f"orc_fill_{kind}")(P(x), first, n, seed)
    return x


SFX = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32", np.dtype(np.uint32): "u32", np.dtype(np.int32): "i32",
       np.dtype(np.uint64): "u64"}
FT = {"f64": C.c_double, "f32": C.c_float}


def orc_stream(kernel: str, a, b, c, scalar=2.0, init_a=1.0):
    """Run one BabelStream kernel of the C oracle in place on numpy arrays (reference argument roles)."""
    L = oracle()
    sfx = SFX[a.dtype]
    n = a.size
    ft = FT[sfx]
    if kernel == "init":
        getattr(L, f"orc_init_{sfx}")(P(a), P(b), P(c), ft(init_a), n)
    elif kernel == "copy":
        getattr(L, f"orc_copy_{sfx}")(P(a), P(b), n)
    elif kernel == "mul":
        getattr(L, f"orc_mul_{sfx}")(P(a), P(b), ft(scalar), n)
    elif kernel == "add":
        getattr(L, f"orc_add_{sfx}")(P(a), P(b), P(c), n)
    elif kernel == "triad":
        getattr(L, f"orc_triad_{sfx}")(P(a), P(b), P(c), ft(scalar), n)
    elif kernel == "nstream":
        getattr(L, f"orc_nstream_{sfx}")(P(a), P(b), P(c), ft(scalar), n)
    else:
        raise ValueError(kernel)


KERNEL_ID = {"init": 0, "copy": 1, "mul": 2, "add": 3, "triad": 4, "nstream": 5, "dot": 6}


def ref_stream(kernel: str, a, b, c, acc=1, init_a=1.0):
    L = ref()
    dtype = 1 if a.dtype == np.float64 else 0
    rc = L.ref_babelstream_run(acc, KERNEL_ID[kernel], dtype, P(a), P(b), P(c), init_a, a.size)
    assert rc == 0, rc


def orc_reduce(x: np.ndarray, block_count: int, block_size: int, iterator: int):
    out = np.zeros(1, dtype=x.dtype)
    rc = getattr(oracle(), f"orc_reduce_{SFX[x.dtype]}")(P(x), x.size, block_count, block_size, iterator, P(out))
    assert rc == 0, rc
    return out[0]


def ref_reduce(x: np.ndarray, acc: int):
    out = np.zeros(1, dtype=x.dtype)
    rc = getattr(ref(), f"ref_reduce_{SFX[x.dtype]}")(acc, P(x), x.size, P(out), None)
    assert rc == 0, rc
    return out[0]


def orc_heat_run(u: np.ndarray, step_first: int, steps: int, dx, dy, dt) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.float64).copy()
    ny, nx = u.shape[0] - 2, u.shape[1] - 2
    rc = oracle().orc_heat2d_run(P(u), ny, nx, step_first, steps, dx, dy, dt)
    assert rc == 0
    return u


def heat_params(ny: int, nx: int):
    """dx, dy as the driver (heatEquation2D.cpp:62-63); dt = 0.2*min(dx^2,dy^2)."""
    dx = 1.0 / (nx + 1)
    dy = 1.0 / (ny + 1)
    dt = 0.2 * min(dx * dx, dy * dy)
    return dx, dy, dt
