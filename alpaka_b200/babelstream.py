"""BabelStream kernels through the C ABI -- the host-side mirror of the reference driver's launches
(reference: benchmarks/babelstream/src/babelStreamMainTest.cpp:305-339 for Init/Copy/Mult/Add/Triad,
:372-405 for Dot). Argument order and meaning follow the reference functors: copy(a -> b), mul(b = s*a),
add(c = a+b), triad(c = a + s*b), nstream(a += b + s*c), dot(a, b)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import B200Error, check
from .runtime import Buf, Queue, alloc_buf, memcpy

SCALAR = 2.0  # scalarVal, babelStreamCommon.hpp:31
INIT_A = 1.0  # valA, babelStreamCommon.hpp:42
MIN_ARRAY_SIZE = 1024 * 128  # minArrSize, babelStreamCommon.hpp:28

# bytes moved per element in units of sizeof(T) (the reference books Init as 2, :415)
ARRAYS_MOVED = {"init": 3, "copy": 2, "mul": 2, "add": 3, "triad": 3, "nstream": 4, "dot": 2}


def _sfx(dtype) -> str:
    dt = np.dtype(dtype)
    if dt == np.float64:
        return "f64"
    if dt == np.float32:
        return "f32"
    raise B200Error(-1, f"BabelStream supports float32/float64, not {dt}")


def _same(n, *bufs):
    dt = bufs[0].dtype
    for b in bufs:
        if b.dtype != dt:
            raise B200Error(-1, "BabelStream buffers must share one element type")
        if len(b.extent) != 1 or b.extent[0] < n:
            raise B200Error(-1, "BabelStream buffers must be 1-D with at least n elements")
    return _sfx(dt)


def _scalar(sfx, v):
    return C.c_double(v) if sfx == "f64" else C.c_float(v)


def init(queue: Queue, a: Buf, b: Buf, c: Buf, init_a: float = INIT_A, n: int | None = None) -> None:
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b, c)
    check(getattr(_lib.load(), f"b200_stream_init_{sfx}")(queue.handle, a.ptr, b.ptr, c.ptr, _scalar(sfx, init_a), n))
    queue._after_enqueue()


def copy(queue: Queue, a: Buf, b: Buf, n: int | None = None) -> None:
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b)
    check(getattr(_lib.load(), f"b200_stream_copy_{sfx}")(queue.handle, a.ptr, b.ptr, n))
    queue._after_enqueue()


def mul(queue: Queue, a: Buf, b: Buf, scalar: float = SCALAR, n: int | None = None) -> None:
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b)
    check(getattr(_lib.load(), f"b200_stream_mul_{sfx}")(queue.handle, a.ptr, b.ptr, _scalar(sfx, scalar), n))
    queue._after_enqueue()


def add(queue: Queue, a: Buf, b: Buf, c: Buf, n: int | None = None) -> None:
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b, c)
    check(getattr(_lib.load(), f"b200_stream_add_{sfx}")(queue.handle, a.ptr, b.ptr, c.ptr, n))
    queue._after_enqueue()


def triad(queue: Queue, a: Buf, b: Buf, c: Buf, scalar: float = SCALAR, n: int | None = None) -> None:
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b, c)
    check(getattr(_lib.load(), f"b200_stream_triad_{sfx}")(queue.handle, a.ptr, b.ptr, c.ptr, _scalar(sfx, scalar), n))
    queue._after_enqueue()


def nstream(queue: Queue, a: Buf, b: Buf, c: Buf, scalar: float = SCALAR, n: int | None = None) -> None:
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b, c)
    check(getattr(_lib.load(), f"b200_stream_nstream_{sfx}")(queue.handle, a.ptr, b.ptr, c.ptr, _scalar(sfx, scalar), n))
    queue._after_enqueue()


def dot_async(queue: Queue, a: Buf, b: Buf, out: Buf, n: int | None = None) -> None:
    """Enqueue the single-pass Dot; the scalar lands in out[0] on the device."""
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b)
    if out.dtype != a.dtype:
        raise B200Error(-1, "dot: result buffer must have the input element type")
    check(getattr(_lib.load(), f"b200_dot_{sfx}")(queue.handle, a.ptr, b.ptr, n, out.ptr, queue.reduce_scratch()))
    queue._after_enqueue()


def dot(queue: Queue, a: Buf, b: Buf, n: int | None = None):
    """Dot product returned to the host (the reference copies 256 block sums back and folds them on the host,
    babelStreamMainTest.cpp:399-403; here one scalar comes back)."""
    out = alloc_buf(queue.dev, a.dtype, 1, queue)
    dot_async(queue, a, b, out, n)
    host = np.empty(1, dtype=a.dtype)
    memcpy(queue, host, out)
    queue.wait()
    out.free()
    return host[0]


def dot_partials(queue: Queue, a: Buf, b: Buf, partials: Buf, n: int | None = None) -> None:
    """Reference-shaped Dot: fills `partials` (e.g. 256 entries) so that their sum is the dot product."""
    n = a.extent[0] if n is None else n
    sfx = _same(n, a, b)
    check(
        getattr(_lib.load(), f"b200_dot_partials_{sfx}")(
            queue.handle, a.ptr, b.ptr, n, partials.ptr, partials.extent[0], queue.reduce_scratch()
        )
    )
    queue._after_enqueue()


class TriadHostPipeline:
    """End-to-end Triad on HOST arrays: c_host = a_host + scalar * b_host, with the host<->device copies inside.

    The arrays are cut into chunks; chunk k runs on stream k % depth as H2D(a), H2D(b), Triad kernel, D2H(c), so the
    two PCIe directions and the kernel overlap across chunks. Host arrays should be pinned (HostBuf) for the copies
    to be asynchronous. This is what bench.py times as `e2e`."""

    def __init__(self, dev, dtype, chunk_elems: int = 1 << 23, depth: int = 4):
        self.dev, self.dtype, self.chunk, self.depth = dev, np.dtype(dtype), int(chunk_elems), int(depth)
        self.queues = [Queue(dev) for _ in range(depth)]
        self.slots = [tuple(alloc_buf(dev, dtype, self.chunk, q) for _ in range(3)) for q in self.queues]

    def run(self, a_host: np.ndarray, b_host: np.ndarray, c_host: np.ndarray, scalar: float = SCALAR) -> tuple[int, int]:
        """Returns (h2d_bytes, d2h_bytes) moved."""
        lib = _lib.load()
        n = a_host.size
        sfx = _sfx(self.dtype)
        fn = getattr(lib, f"b200_stream_triad_{sfx}")
        item = self.dtype.itemsize
        pa, pb, pc = a_host.ctypes.data, b_host.ctypes.data, c_host.ctypes.data
        k = 0
        for start in range(0, n, self.chunk):
            m = min(self.chunk, n - start)
            q = self.queues[k % self.depth]
            da, db, dc = self.slots[k % self.depth]
            off = start * item
            check(lib.b200_memcpy_async(self.dev.idx, da.ptr, pa + off, m * item, 1, q.handle))
            check(lib.b200_memcpy_async(self.dev.idx, db.ptr, pb + off, m * item, 1, q.handle))
            check(fn(q.handle, da.ptr, db.ptr, dc.ptr, _scalar(sfx, scalar), m))
            check(lib.b200_memcpy_async(self.dev.idx, pc + off, dc.ptr, m * item, 2, q.handle))
            k += 1
        for q in self.queues:
            q.wait()
        return 2 * n * item, n * item

    def close(self):
        for slot in self.slots:
            for b in slot:
                b.free()
        for q in self.queues:
            q.close()
