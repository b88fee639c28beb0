"""Host-side mirror of alpaka's platform / device / queue / event / buffer API over the C ABI.

Python is only the test and benchmark driver here; the product's host side is the C++20 header layer in
include/alpaka/ which binds the same C ABI. Names follow the reference's free functions (reference files cited per
item; paths relative to the reference root):

    Platform, get_dev_by_idx, get_dev_count      platform/Traits.hpp:54-81
    Dev.name / mem_bytes / free_mem_bytes         dev/Traits.hpp:56-126
    Queue (blocking | non-blocking), wait, empty  queue/Traits.hpp:46-70, wait/Traits.hpp:33-49
    Event, enqueue(queue, event), is_complete     event/Traits.hpp
    alloc_buf / alloc_async_buf / alloc_mapped_buf, Buf, memcpy, memset, get_pitches_in_bytes
                                                  mem/buf/Traits.hpp:63-137, mem/view/Traits.hpp:207-310
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import B200Error, check

COPY_H2H, COPY_H2D, COPY_D2H, COPY_D2D, COPY_DEFAULT = 0, 1, 2, 3, 4
REDUCE_SCRATCH_BYTES = 65536


class Platform:
    """alpaka::Platform<Acc> for the B200 accelerator (an empty value type in the reference,
    platform/PlatformUniformCudaHipRt.hpp:28-39)."""

    def get_dev_count(self) -> int:
        n = C.c_int(0)
        check(_lib.load().b200_device_count(C.byref(n)))
        return n.value

    def get_dev_by_idx(self, idx: int) -> "Dev":
        n = self.get_dev_count()
        if not 0 <= idx < n:
            # platform/PlatformUniformCudaHipRt.hpp:75-83 throws for an out-of-range index
            raise B200Error(-1, f"Unable to return device handle for device {idx}. There are only {n} devices!")
        return Dev(idx)

    def enable_peer_access(self) -> int:
        n = C.c_int(0)
        check(_lib.load().b200_enable_peer_all(C.byref(n)))
        return n.value


def get_dev_by_idx(platform: Platform, idx: int) -> "Dev":
    return platform.get_dev_by_idx(idx)


def get_dev_count(platform: Platform) -> int:
    return platform.get_dev_count()


class Dev:
    """alpaka::DevCudaRt equivalent: a cheap copyable handle = device ordinal (dev/DevUniformCudaHipRt.hpp:55-107)."""

    def __init__(self, idx: int):
        self.idx = int(idx)

    def __eq__(self, other):
        return isinstance(other, Dev) and other.idx == self.idx

    def __hash__(self):
        return hash(("b200dev", self.idx))

    def props(self) -> _lib.DeviceProps:
        p = _lib.DeviceProps()
        check(_lib.load().b200_device_props_get(self.idx, C.byref(p)))
        return p

    @property
    def name(self) -> str:
        return self.props().name.decode()

    @property
    def mem_bytes(self) -> int:
        return int(self.props().total_global_mem)

    @property
    def free_mem_bytes(self) -> int:
        f, t = C.c_uint64(), C.c_uint64()
        check(_lib.load().b200_device_mem_info(self.idx, C.byref(f), C.byref(t)))
        return f.value

    @property
    def multi_processor_count(self) -> int:
        return int(self.props().multi_processor_count)

    def wait(self) -> None:
        check(_lib.load().b200_device_sync(self.idx))

    def pool_stats(self) -> tuple[int, int]:
        r, u = C.c_uint64(), C.c_uint64()
        check(_lib.load().b200_pool_stats(self.idx, C.byref(r), C.byref(u)))
        return r.value, u.value


class Queue:
    """alpaka::Queue<Acc, Blocking|NonBlocking>: one cudaStreamNonBlocking stream; a blocking queue synchronises after
    every enqueue (queue/cuda_hip/QueueUniformCudaHipRt.hpp:40-127, kernel/TaskKernelGpuUniformCudaHipRt.hpp:284-290)."""

    def __init__(self, dev: Dev, blocking: bool = False, *, native_handle: Optional[int] = None):
        self.dev = dev
        self.blocking = bool(blocking)
        self._owns = native_handle is None
        if native_handle is None:
            h = C.c_void_p()
            check(_lib.load().b200_stream_create(dev.idx, C.byref(h)))
            self.handle = h.value
        else:
            # adopt an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); never destroyed here
            self.handle = native_handle
        self._scratch: Optional[int] = None

    def wait(self) -> None:
        check(_lib.load().b200_stream_sync(self.handle))

    def empty(self) -> bool:
        v = C.c_int(0)
        check(_lib.load().b200_stream_query(self.handle, C.byref(v)))
        return bool(v.value)

    def _after_enqueue(self) -> None:
        if self.blocking:
            self.wait()

    def reduce_scratch(self) -> int:
        """Per-queue scratch of the single-pass grid reductions (zeroed once)."""
        if self._scratch is None:
            p = C.c_void_p()
            lib = _lib.load()
            check(lib.b200_malloc_async(self.dev.idx, self.handle, REDUCE_SCRATCH_BYTES, C.byref(p)))
            check(lib.b200_memset_async(self.dev.idx, p, 0, REDUCE_SCRATCH_BYTES, self.handle))
            self._scratch = p.value
        return self._scratch

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib = _lib.load()
            if self._scratch is not None:
                lib.b200_free_async(self.dev.idx, self.handle, self._scratch)
                self._scratch = None
            if self._owns:
                lib.b200_stream_destroy(self.dev.idx, self.handle)
            else:
                lib.b200_stream_sync(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Event:
    """alpaka::Event<Queue> (event/EventUniformCudaHipRt.hpp:26-112). timing=True is a benchmark extension; the
    reference's events are created with cudaEventDisableTiming."""

    def __init__(self, dev: Dev, timing: bool = False):
        self.dev = dev
        h = C.c_void_p()
        check(_lib.load().b200_event_create(dev.idx, int(timing), C.byref(h)))
        self.handle = h.value

    def record(self, queue: Queue) -> None:
        check(_lib.load().b200_event_record(self.handle, queue.handle))
        queue._after_enqueue()

    def is_complete(self) -> bool:
        v = C.c_int(0)
        check(_lib.load().b200_event_query(self.handle, C.byref(v)))
        return bool(v.value)

    def wait(self) -> None:
        check(_lib.load().b200_event_sync(self.handle))

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float(0)
        check(_lib.load().b200_event_elapsed_ms(self.handle, stop.handle, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().b200_event_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def enqueue(queue: Queue, event: Event) -> None:
    event.record(queue)


def wait(x, awaited: Optional[Event] = None) -> None:
    """alpaka::wait(x) / alpaka::wait(waiter, awaited) (wait/Traits.hpp:33-49)."""
    if awaited is None:
        x.wait()
    elif isinstance(x, Queue):
        check(_lib.load().b200_stream_wait_event(x.handle, awaited.handle))
    elif isinstance(x, Dev):
        check(_lib.load().b200_device_wait_event(x.idx, awaited.handle))
    else:
        raise TypeError("waiter must be a Queue or a Dev")


Extent = Union[int, Sequence[int]]


def _extent_tuple(extent: Extent) -> tuple[int, ...]:
    if isinstance(extent, (int, np.integer)):
        return (int(extent),)
    return tuple(int(e) for e in extent)


class Buf:
    """alpaka::Buf<Dev, T, Dim, Idx> on the device: pointer + extent + row pitch
    (mem/buf/BufUniformCudaHipRt.hpp:53-88). 1-D and 2-D only (the hot path uses nothing else)."""

    def __init__(self, dev: Dev, dtype, extent: Extent, queue: Optional[Queue] = None, ipc: bool = False, *,
                 native_ptr: Optional[int] = None, pitch_bytes: Optional[int] = None):
        self.dev = dev
        self.dtype = np.dtype(dtype)
        self.extent = _extent_tuple(extent)
        if len(self.extent) not in (1, 2):
            raise B200Error(-1, "only 1-D and 2-D buffers are supported")
        self._queue = queue
        self._ipc = ipc
        self._view = native_ptr is not None
        lib = _lib.load()
        if native_ptr is not None:
            # alpaka::createView(dev, ptr, extent[, pitch]) (mem/view/Traits.hpp:434-482): non-owning
            width_bytes = self.extent[-1] * self.dtype.itemsize
            self.pitch_bytes = int(pitch_bytes) if pitch_bytes is not None else width_bytes
            self.nbytes = self.pitch_bytes * self.extent[0] if len(self.extent) == 2 else width_bytes
            self.ptr = int(native_ptr)
            return
        p = C.c_void_p()
        width_bytes = self.extent[-1] * self.dtype.itemsize
        if len(self.extent) == 2:
            self.pitch_bytes = int(lib.b200_pitch_for_width(width_bytes))
            nbytes = self.pitch_bytes * self.extent[0]
        else:
            self.pitch_bytes = width_bytes
            nbytes = width_bytes
        self.nbytes = nbytes
        if ipc:
            check(lib.b200_malloc_device(dev.idx, nbytes, C.byref(p)))
        else:
            check(lib.b200_malloc_async(dev.idx, queue.handle if queue else None, nbytes, C.byref(p)))
        self.ptr = p.value or 0

    def get_pitches_in_bytes(self) -> tuple[int, ...]:
        """getPitchesInBytes: [Dim-1] == sizeof(T), [Dim-2] == row pitch (BufUniformCudaHipRt.hpp:189-199)."""
        if len(self.extent) == 2:
            return (self.pitch_bytes, self.dtype.itemsize)
        return (self.dtype.itemsize,)

    def free(self) -> None:
        if getattr(self, "_view", False):
            self.ptr = 0
            return
        if getattr(self, "ptr", 0):
            lib = _lib.load()
            if self._ipc:
                lib.b200_free_device(self.dev.idx, self.ptr)
            elif self._queue is not None:
                lib.b200_free_async(self.dev.idx, self._queue.handle, self.ptr)
            else:
                # allocated without a queue: the pool free would be ordered on the legacy NULL stream, which the
                # (non-blocking) queue streams do not synchronise with. cudaFree semantics instead: wait for the device
                # first, as the C++ allocBuf deleter does (include/alpaka/b200/Mem.hpp), then release.
                lib.b200_device_sync(self.dev.idx)
                lib.b200_free_async(self.dev.idx, None, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class HostBuf:
    """Pinned host buffer (alpaka::allocMappedBuf, mem/buf/BufUniformCudaHipRt.hpp:338-360) exposed as numpy."""

    def __init__(self, dtype, extent: Extent):
        self.dtype = np.dtype(dtype)
        self.extent = _extent_tuple(extent)
        n = int(np.prod(self.extent)) if self.extent else 1
        self.nbytes = n * self.dtype.itemsize
        p = C.c_void_p()
        check(_lib.load().b200_host_alloc_pinned(max(self.nbytes, 1), C.byref(p)))
        self.ptr = p.value
        raw = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.extent)
        self.pitch_bytes = self.extent[-1] * self.dtype.itemsize if self.extent else self.dtype.itemsize

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.array = None
                _lib.load().b200_host_free_pinned(self.ptr)
                self.ptr = None
        except Exception:
            pass


def alloc_buf(dev: Dev, dtype, extent: Extent, queue: Optional[Queue] = None) -> Buf:
    """alpaka::allocBuf<T, Idx>(dev, extent): served from the device's stream-ordered pool."""
    return Buf(dev, dtype, extent, queue)


def alloc_async_buf(queue: Queue, dtype, extent: Extent) -> Buf:
    """alpaka::allocAsyncBuf<T, Idx>(queue, extent) (mem/buf/Traits.hpp:78-82)."""
    return Buf(queue.dev, dtype, extent, queue)


def create_view(dev: Dev, native_ptr: int, dtype, extent: Extent, pitch_bytes: Optional[int] = None) -> Buf:
    """alpaka::createView(dev, pointer, extent[, pitch]): a non-owning Buf over existing device memory."""
    return Buf(dev, dtype, extent, native_ptr=native_ptr, pitch_bytes=pitch_bytes)


def alloc_mapped_buf(dtype, extent: Extent) -> HostBuf:
    return HostBuf(dtype, extent)


def _host_view(x):
    if isinstance(x, HostBuf):
        return x.array
    return x


def memcpy(queue: Queue, dst, src, extent: Optional[Extent] = None) -> None:
    """alpaka::memcpy(queue, dst, src[, extent]) between device Buf and host numpy/HostBuf, or Buf to Buf.
    Element types must match and dims must agree (mem/view/Traits.hpp:262-278)."""
    lib = _lib.load()
    d_dev, s_dev = isinstance(dst, Buf), isinstance(src, Buf)
    dh, sh = _host_view(dst), _host_view(src)
    d_dtype = dst.dtype if d_dev else dh.dtype
    s_dtype = src.dtype if s_dev else sh.dtype
    if d_dtype != s_dtype:
        raise B200Error(-1, "memcpy: element types of destination and source differ")
    d_ext = dst.extent if d_dev else tuple(dh.shape)
    s_ext = src.extent if s_dev else tuple(sh.shape)
    ext = _extent_tuple(extent) if extent is not None else d_ext
    if not (len(d_ext) == len(s_ext) == len(ext)):
        raise B200Error(-1, "memcpy: dimensionality of destination, source and extent differ")
    if any(e > d or e > s for e, d, s in zip(ext, d_ext, s_ext)):
        raise B200Error(-1, "memcpy: extent exceeds destination or source")
    for h in (dh if not d_dev else None, sh if not s_dev else None):
        if h is not None and not h.flags["C_CONTIGUOUS"]:
            raise B200Error(-1, "memcpy: host arrays must be C-contiguous")
    kind = {(True, True): COPY_D2D, (True, False): COPY_H2D, (False, True): COPY_D2H, (False, False): COPY_H2H}[(d_dev, s_dev)]
    dptr = dst.ptr if d_dev else dh.ctypes.data
    sptr = src.ptr if s_dev else sh.ctypes.data
    item = d_dtype.itemsize
    if len(ext) == 1:
        check(lib.b200_memcpy_async(queue.dev.idx, dptr, sptr, ext[0] * item, kind, queue.handle))
    else:
        dpitch = dst.pitch_bytes if d_dev else d_ext[1] * item
        spitch = src.pitch_bytes if s_dev else s_ext[1] * item
        check(lib.b200_memcpy2d_async(queue.dev.idx, dptr, dpitch, sptr, spitch, ext[1] * item, ext[0], kind, queue.handle))
    if not (d_dev and s_dev):
        # pageable host memory: cudaMemcpyAsync may return before the transfer; numpy owners expect completion
        if not isinstance(dst, HostBuf) and not isinstance(src, HostBuf):
            queue.wait()
    queue._after_enqueue()


def memset(queue: Queue, buf: Buf, byte: int, extent: Optional[Extent] = None) -> None:
    """alpaka::memset(queue, view, byte[, extent]) (mem/view/Traits.hpp:207-245)."""
    lib = _lib.load()
    ext = _extent_tuple(extent) if extent is not None else buf.extent
    item = buf.dtype.itemsize
    if len(ext) == 1:
        check(lib.b200_memset_async(buf.dev.idx, buf.ptr, byte, ext[0] * item, queue.handle))
    else:
        check(lib.b200_memset2d_async(buf.dev.idx, buf.ptr, buf.pitch_bytes, byte, ext[1] * item, ext[0], queue.handle))
    queue._after_enqueue()


def launch_count() -> int:
    return int(_lib.load().b200_launch_count())


def tune_set(key: str, value: int) -> None:
    check(_lib.load().b200_tune_set(key.encode(), int(value)))
