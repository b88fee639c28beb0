"""Runtime half of the C ABI on a real device: buffers, copies, queues, events (semantics pinned by the reference's
unit tests: test/unit/mem/buf/src/BufTest.cpp, test/unit/queue/src/QueueTest.cpp, test/unit/event/src/EventTest.cpp)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_props(gpu):
    ab, dev, queue = gpu
    p = dev.props()
    assert p.cc_major == 10, "built for sm_100a only"
    assert p.multi_processor_count >= 100 and p.warp_size == 32
    assert dev.free_mem_bytes <= dev.mem_bytes
    assert "NVIDIA" in dev.name


def test_buf_roundtrip_1d_2d_and_pitch(gpu):
    ab, dev, queue = gpu
    x = np.arange(10007, dtype=np.float64)
    b = ab.alloc_buf(dev, np.float64, x.size, queue)
    ab.memcpy(queue, b, x)
    y = np.empty_like(x)
    ab.memcpy(queue, y, b)
    queue.wait()
    assert (x == y).all()
    m = np.arange(37 * 53, dtype=np.float64).reshape(37, 53)
    b2 = ab.alloc_buf(dev, np.float64, m.shape, queue)
    pitches = b2.get_pitches_in_bytes()
    assert pitches[1] == 8 and pitches[0] % 128 == 0 and pitches[0] >= 53 * 8
    ab.memcpy(queue, b2, m)
    m2 = np.zeros_like(m)
    ab.memcpy(queue, m2, b2)
    queue.wait()
    assert (m == m2).all()
    # sub-extent copy (test/unit/mem/view BufSlicing)
    sub = np.zeros((5, 7))
    ab.memcpy(queue, sub, b2, (5, 7))
    queue.wait()
    assert (sub == m[:5, :7]).all()


def test_zero_size_buffer_and_memset(gpu):
    ab, dev, queue = gpu
    z = ab.alloc_buf(dev, np.float32, 0, queue)
    assert z.ptr == 0
    b = ab.alloc_buf(dev, np.uint8, 4096, queue)
    ab.memset(queue, b, 0x5A)
    out = np.zeros(4096, dtype=np.uint8)
    ab.memcpy(queue, out, b)
    queue.wait()
    assert (out == 0x5A).all()


def test_memcpy_checks(gpu):
    ab, dev, queue = gpu
    b = ab.alloc_buf(dev, np.float64, 16, queue)
    with pytest.raises(ab.B200Error):
        ab.memcpy(queue, b, np.zeros(16, dtype=np.float32))
    with pytest.raises(ab.B200Error):
        ab.memcpy(queue, b, np.zeros((4, 4)))
    with pytest.raises(ab.B200Error):
        ab.memcpy(queue, b, np.zeros(8), 16)


def test_pool_reuses_memory(gpu):
    ab, dev, queue = gpu
    b = ab.alloc_buf(dev, np.uint8, 64 << 20, queue)
    queue.wait()
    reserved0, used0 = dev.pool_stats()
    b.free()
    queue.wait()
    reserved1, used1 = dev.pool_stats()
    assert used1 < used0 and reserved1 >= 64 << 20  # freed to the pool, not to the OS
    b2 = ab.alloc_buf(dev, np.uint8, 64 << 20, queue)
    queue.wait()
    assert dev.pool_stats()[0] == reserved1
    b2.free()


def test_event_ordering_and_timing(gpu):
    ab, dev, queue = gpu
    q2 = ab.Queue(dev)
    a = ab.alloc_buf(dev, np.float64, 1 << 24, queue)
    c = ab.alloc_buf(dev, np.float64, 1 << 24, queue)
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
    ab.enqueue(queue, e0)
    ab.babelstream.init(queue, a, c, c, 2.0)
    ab.enqueue(queue, e1)
    ab.wait(q2, e1)  # q2 waits for queue's event
    out = ab.alloc_buf(dev, np.float64, 1 << 24, q2)
    ab.babelstream.copy(q2, a, out)
    q2.wait()
    assert e1.is_complete()
    assert e0.elapsed_ms(e1) > 0
    h = np.empty(1 << 24)
    ab.memcpy(q2, h, out)
    q2.wait()
    assert (h == 2.0).all()
    q2.close()


def test_pinned_host_buffer(gpu):
    ab, dev, queue = gpu
    hb = ab.alloc_mapped_buf(np.float32, 1 << 16)
    hb.array[:] = np.arange(1 << 16, dtype=np.float32)
    b = ab.alloc_buf(dev, np.float32, 1 << 16, queue)
    ab.memcpy(queue, b, hb)
    hb2 = ab.alloc_mapped_buf(np.float32, 1 << 16)
    ab.memcpy(queue, hb2, b)
    queue.wait()
    assert (hb2.array == hb.array).all()


def test_launch_counter_and_error_string(gpu):
    ab, dev, queue = gpu
    from alpaka_b200 import _lib

    n0 = ab.runtime.launch_count()
    a = ab.alloc_buf(dev, np.float64, 1024, queue)
    ab.babelstream.copy(queue, a, a)
    assert ab.runtime.launch_count() == n0 + 1
    lib = _lib.load()
    rc = lib.b200_stream_copy_f64(queue.handle, None, None, 5)
    assert rc == -1 and b"B200_EINVAL" in lib.b200_last_error_string()
