"""ctypes binding of libalpaka_b200.so (C ABI: include/b200/b200.h).

The library is the product; this module only loads it and declares signatures. There is NO CPU fallback:
if the shared library is missing the import fails loudly, and without a CUDA device every entry that needs
one returns an error that `check()` raises as `B200Error`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libalpaka_b200.so")


class B200Error(RuntimeError):
    """Raised for any non-zero return code of the C ABI (mirrors the std::runtime_error the reference throws,
    reference: include/alpaka/core/UniformCudaHip.hpp:23-112)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[b200 rc={code}] {message}")
        self.code = code


class DeviceProps(C.Structure):
    _fields_ = [
        ("name", C.c_char * 256),
        ("cc_major", C.c_int32),
        ("cc_minor", C.c_int32),
        ("multi_processor_count", C.c_int32),
        ("max_grid_dim", C.c_int32 * 3),
        ("max_block_dim", C.c_int32 * 3),
        ("max_threads_per_block", C.c_int32),
        ("warp_size", C.c_int32),
        ("shared_mem_per_block", C.c_uint64),
        ("shared_mem_per_block_optin", C.c_uint64),
        ("total_global_mem", C.c_uint64),
        ("free_global_mem", C.c_uint64),
        ("l2_cache_bytes", C.c_int32),
        ("memory_pools_supported", C.c_int32),
    ]


class FuncAttributes(C.Structure):
    _fields_ = [
        ("max_threads_per_block", C.c_int32),
        ("num_regs", C.c_int32),
        ("shared_size_bytes", C.c_uint64),
        ("const_size_bytes", C.c_uint64),
        ("local_size_bytes", C.c_uint64),
        ("max_dynamic_shared_size_bytes", C.c_int32),
        ("ptx_version", C.c_int32),
        ("binary_version", C.c_int32),
    ]


class AccDevProps(C.Structure):
    _fields_ = [
        ("multi_processor_count", C.c_uint64),
        ("grid_block_extent_max", C.c_uint64 * 4),
        ("grid_block_count_max", C.c_uint64),
        ("block_thread_extent_max", C.c_uint64 * 4),
        ("block_thread_count_max", C.c_uint64),
        ("thread_elem_extent_max", C.c_uint64 * 4),
        ("thread_elem_count_max", C.c_uint64),
        ("shared_mem_size_bytes", C.c_uint64),
        ("global_mem_size_bytes", C.c_uint64),
    ]


class Heat2dHalo(C.Structure):
    """b200_heat2d_halo: peer pointers and flag words of the fused halo exchange (side order top, bottom, left, right)."""

    _fields_ = [
        ("peer_u", (C.c_void_p * 2) * 4),
        ("peer_flag", C.c_void_p * 4),
        ("my_flags", C.c_void_p),
    ]


class Heat2dWalkPlan(C.Structure):
    """b200_heat2d_walk_plan: the host-side decomposition of a walker launch (b200_heat2d_walk_plan_query)."""

    _fields_ = [
        ("window_columns", C.c_uint32), ("n_windows", C.c_uint32), ("n_edge_right", C.c_uint32),
        ("n_front", C.c_uint32), ("front_is_strip", C.c_uint32),
        ("front_y0", C.c_int32 * 2), ("front_y1", C.c_int32 * 2),
        ("interior_y0", C.c_int32), ("interior_y1", C.c_int32), ("segment_rows", C.c_int32),
        ("n_segments", C.c_uint32), ("n_walkers", C.c_uint32), ("split", C.c_int32),
    ]


class Exchange(C.Structure):
    """b200_exchange: every rank's exchange buffer as seen from this device (fused Dot / reduce over several GPUs)."""

    _fields_ = [("base", C.c_void_p * 16), ("world", C.c_uint32), ("rank", C.c_uint32)]


EXCHANGE_BYTES = 512

_vp, _u64, _u32, _i, _sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
_f64, _f32 = C.c_double, C.c_float
_P = C.POINTER

# name -> argtypes (restype is int unless listed in _RESTYPES). Must list EVERY symbol include/b200/b200.h declares;
# tests/test_abi.py parses the header and checks this table and the .so against it.
SIGNATURES: dict[str, list] = {
    "b200_abi_version": [],
    "b200_last_error_string": [],
    "b200_error_name": [_i],
    "b200_device_count": [_P(_i)],
    "b200_device_props_get": [_i, _P(DeviceProps)],
    "b200_device_mem_info": [_i, _P(_u64), _P(_u64)],
    "b200_device_sync": [_i],
    "b200_device_reset": [_i],
    "b200_enable_peer_all": [_P(_i)],
    "b200_stream_create": [_i, _P(_vp)],
    "b200_stream_destroy": [_i, _vp],
    "b200_stream_sync": [_vp],
    "b200_stream_query": [_vp, _P(_i)],
    "b200_launch_host_func": [_vp, _vp, _vp],
    "b200_event_create": [_i, _i, _P(_vp)],
    "b200_event_destroy": [_vp],
    "b200_event_record": [_vp, _vp],
    "b200_event_query": [_vp, _P(_i)],
    "b200_event_sync": [_vp],
    "b200_stream_wait_event": [_vp, _vp],
    "b200_device_wait_event": [_i, _vp],
    "b200_event_elapsed_ms": [_vp, _vp, _P(_f32)],
    "b200_malloc_async": [_i, _vp, _sz, _P(_vp)],
    "b200_free_async": [_i, _vp, _vp],
    "b200_malloc_pitched_async": [_i, _vp, _sz, _sz, _P(_vp), _P(_sz)],
    "b200_pitch_for_width": [_sz],
    "b200_malloc_device": [_i, _sz, _P(_vp)],
    "b200_free_device": [_i, _vp],
    "b200_host_alloc_pinned": [_sz, _P(_vp)],
    "b200_host_free_pinned": [_vp],
    "b200_host_register": [_vp, _sz],
    "b200_host_unregister": [_vp],
    "b200_pool_stats": [_i, _P(_u64), _P(_u64)],
    "b200_pool_trim": [_i, _sz],
    "b200_memcpy_async": [_i, _vp, _vp, _sz, _i, _vp],
    "b200_memcpy2d_async": [_i, _vp, _sz, _vp, _sz, _sz, _sz, _i, _vp],
    "b200_memcpy_peer_async": [_vp, _i, _vp, _i, _sz, _vp],
    "b200_memset_async": [_i, _vp, _i, _sz, _vp],
    "b200_memset2d_async": [_i, _vp, _sz, _i, _sz, _sz, _vp],
    "b200_ipc_get_mem_handle": [_i, _vp, C.c_char_p],
    "b200_ipc_open_mem_handle": [_i, C.c_char_p, _P(_vp)],
    "b200_ipc_close_mem_handle": [_i, _vp],
    "b200_ipc_event_create": [_i, _P(_vp), C.c_char_p],
    "b200_ipc_event_open": [_i, C.c_char_p, _P(_vp)],
    "b200_func_attributes_get": [_i, _vp, _P(FuncAttributes)],
    "b200_launch": [_i, _vp, _P(_u32), _P(_u32), _sz, _vp, _P(_vp)],
    "b200_mem_range": [_vp, _P(_vp), _P(_sz)],
    "b200_acc_dev_props_get": [_i, _i, _P(AccDevProps)],
    "b200_subdivide_grid_elems": [_i, _P(_u64), _P(_u64), _P(AccDevProps), _u64, _i, _i, _P(_u64), _P(_u64), _P(_u64)],
    "b200_is_valid_work_div": [_i, _P(_u64), _P(_u64), _P(_u64), _P(AccDevProps), _u64, _P(_i)],
    "b200_stream_init_f64": [_vp, _vp, _vp, _vp, _f64, _u64],
    "b200_stream_copy_f64": [_vp, _vp, _vp, _u64],
    "b200_stream_mul_f64": [_vp, _vp, _vp, _f64, _u64],
    "b200_stream_add_f64": [_vp, _vp, _vp, _vp, _u64],
    "b200_stream_triad_f64": [_vp, _vp, _vp, _vp, _f64, _u64],
    "b200_stream_nstream_f64": [_vp, _vp, _vp, _vp, _f64, _u64],
    "b200_stream_init_f32": [_vp, _vp, _vp, _vp, _f32, _u64],
    "b200_stream_copy_f32": [_vp, _vp, _vp, _u64],
    "b200_stream_mul_f32": [_vp, _vp, _vp, _f32, _u64],
    "b200_stream_add_f32": [_vp, _vp, _vp, _vp, _u64],
    "b200_stream_triad_f32": [_vp, _vp, _vp, _vp, _f32, _u64],
    "b200_stream_nstream_f32": [_vp, _vp, _vp, _vp, _f32, _u64],
    "b200_dot_f64": [_vp, _vp, _vp, _u64, _vp, _vp],
    "b200_dot_f32": [_vp, _vp, _vp, _u64, _vp, _vp],
    "b200_dot_partials_f64": [_vp, _vp, _vp, _u64, _vp, _u32, _vp],
    "b200_dot_partials_f32": [_vp, _vp, _vp, _u64, _vp, _u32, _vp],
    "b200_reduce_sum_u32": [_vp, _vp, _u64, _vp, _vp],
    "b200_reduce_sum_i32": [_vp, _vp, _u64, _vp, _vp],
    "b200_reduce_sum_u64": [_vp, _vp, _u64, _vp, _vp],
    "b200_reduce_sum_f32": [_vp, _vp, _u64, _vp, _vp],
    "b200_reduce_sum_f64": [_vp, _vp, _u64, _vp, _vp],
    "b200_dot_allranks_f64": [_vp, _vp, _vp, _u64, _vp, _vp, _P(Exchange), _u32],
    "b200_dot_allranks_f32": [_vp, _vp, _vp, _u64, _vp, _vp, _P(Exchange), _u32],
    "b200_reduce_sum_allranks_u32": [_vp, _vp, _u64, _vp, _vp, _P(Exchange), _u32],
    "b200_reduce_sum_allranks_f32": [_vp, _vp, _u64, _vp, _vp, _P(Exchange), _u32],
    "b200_reduce_sum_allranks_f64": [_vp, _vp, _u64, _vp, _vp, _P(Exchange), _u32],
    "b200_exchange_status": [_i, _vp, _P(_u32)],
    "b200_heat2d_plan_create": [_i, _vp, _vp, _sz, _u32, _u32, _vp, _vp, _i, _P(_vp)],
    "b200_heat2d_plan_destroy": [_vp],
    "b200_heat2d_step_f64": [_vp, _vp, _i, _f64, _f64, _f64],
    "b200_heat2d_step2_f64": [_vp, _vp, _i, _f64, _f64, _f64, _f64],
    "b200_heat2d_stepn_f64": [_vp, _vp, _i, _f64, _f64, _i, _vp],
    "b200_heat2d_slab_plan_create": [_i, _vp, _vp, _sz, _u32, _u32, _vp, _vp, _i, _u32, _P(_vp)],
    "b200_heat2d_stepn_halo_f64": [_vp, _vp, _i, _f64, _f64, _i, _vp, _u32],
    "b200_heat2d_tile_plan_create": [_i, _vp, _vp, _sz, _u32, _u32, _vp, _vp, _i, _u32, _P(_vp)],
    "b200_heat2d_stepn_tile_f64": [_vp, _vp, _i, _f64, _f64, _i, _vp, _u32],
    "b200_heat2d_walk_plan_query": [_u32, _u32, _u32, _u32, _i, _i, _i, _vp],
    "b200_heat2d_step2_halo_f64": [_vp, _vp, _i, _f64, _f64, _f64, _f64, _u32],
    "b200_heat2d_step_window_f64": [_vp, _vp, _i, _f64, _f64, _f64, _u32, _u32, _u32, _u32],
    "b200_heat2d_boundary_f64": [_vp, _vp, _i, _f64],
    "b200_heat2d_plan_set_halo": [_vp, _P(Heat2dHalo)],
    "b200_heat2d_step_halo_f64": [_vp, _vp, _i, _f64, _f64, _f64, _u32],
    "b200_heat2d_halo_status": [_vp, _P(_u32)],
    "b200_tune_set": [C.c_char_p, C.c_int64],
    "b200_tune_get": [C.c_char_p, _P(C.c_int64)],
    "b200_launch_count": [],
}
_RESTYPES = {
    "b200_last_error_string": C.c_char_p,
    "b200_error_name": C.c_char_p,
    "b200_pitch_for_width": C.c_size_t,
    "b200_launch_count": C.c_uint64,
}

_lib = None


def load() -> C.CDLL:
    """Load libalpaka_b200.so (once). Raises ImportError with the build command if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C alpaka_b200/csrc`). alpaka_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift; fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.b200_abi_version() != 1:
        raise ImportError("libalpaka_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        lib = load()
        msg = lib.b200_last_error_string()
        name = lib.b200_error_name(rc)
        raise B200Error(rc, (msg or b"").decode() or (name or b"?").decode())
