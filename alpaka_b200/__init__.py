"""alpaka_b200: B200-native implementation of alpaka's data-parallel hot path (BabelStream kernels, example/reduce,
example/heatEquation2D) behind a C ABI (include/b200/b200.h, alpaka_b200/lib/libalpaka_b200.so) with the alpaka-
compatible C++20 header layer in include/alpaka/. This Python package is the thin ctypes host mirror used by
tests/ and bench.py. No CPU fallback: importing it without the built library raises ImportError."""
from . import _lib
from ._lib import B200Error

_lib.load()  # fail loudly if the CUDA library has not been built

from . import babelstream, decomp, heat2d, multi, reduce, runtime, workdiv  # noqa: E402
from .runtime import (  # noqa: E402
    Buf,
    Dev,
    Event,
    HostBuf,
    Platform,
    Queue,
    alloc_async_buf,
    alloc_buf,
    alloc_mapped_buf,
    create_view,
    enqueue,
    get_dev_by_idx,
    get_dev_count,
    memcpy,
    memset,
    wait,
)

__all__ = [
    "B200Error", "Buf", "Dev", "Event", "HostBuf", "Platform", "Queue", "alloc_async_buf", "alloc_buf",
    "alloc_mapped_buf", "babelstream", "create_view", "decomp", "multi", "enqueue", "get_dev_by_idx", "get_dev_count", "heat2d", "memcpy", "memset",
    "reduce", "runtime", "wait", "workdiv",
]
