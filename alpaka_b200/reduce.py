"""example/reduce through the C ABI: sum of a 1-D device buffer in one single-pass launch
(reference: example/reduce/src/reduce.cpp:47-105 issues two ReduceKernel launches and copies one element back)."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import B200Error, check
from .runtime import Buf, Queue, alloc_buf, memcpy

_SFX = {
    np.dtype(np.uint32): "u32",
    np.dtype(np.int32): "i32",
    np.dtype(np.uint64): "u64",
    np.dtype(np.float32): "f32",
    np.dtype(np.float64): "f64",
}


def reduce_sum_async(queue: Queue, source: Buf, out: Buf, n: int | None = None) -> None:
    n = source.extent[0] if n is None else n
    sfx = _SFX.get(source.dtype)
    if sfx is None:
        raise B200Error(-1, f"reduce: unsupported element type {source.dtype}")
    if out.dtype != source.dtype:
        raise B200Error(-1, "reduce: result buffer must have the input element type")
    if len(source.extent) != 1 or n > source.extent[0]:
        raise B200Error(-1, "reduce: source must be 1-D with at least n elements")
    check(getattr(_lib.load(), f"b200_reduce_sum_{sfx}")(queue.handle, source.ptr, n, out.ptr, queue.reduce_scratch()))
    queue._after_enqueue()


def reduce_sum(queue: Queue, source: Buf, n: int | None = None):
    out = alloc_buf(queue.dev, source.dtype, 1, queue)
    reduce_sum_async(queue, source, out, n)
    host = np.empty(1, dtype=source.dtype)
    memcpy(queue, host, out)
    queue.wait()
    out.free()
    return host[0]
