"""Parity of the CUDA BabelStream path (through the C ABI) against the oracle. Bit-exact for the element-wise
kernels, <= 1e-12 (double) / 1e-5 (float) relative for Dot (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import P

pytestmark = pytest.mark.gpu

KERNELS = ["init", "copy", "mul", "add", "triad", "nstream"]
DOT_RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def _run_gpu(ab, queue, kernel, da, db, dc, scalar, init_a, n=None):
    bs = ab.babelstream
    if kernel == "init":
        bs.init(queue, da, db, dc, init_a, n)
    elif kernel == "copy":
        bs.copy(queue, da, db, n)
    elif kernel == "mul":
        bs.mul(queue, da, db, scalar, n)
    elif kernel == "add":
        bs.add(queue, da, db, dc, n)
    elif kernel == "triad":
        bs.triad(queue, da, db, dc, scalar, n)
    elif kernel == "nstream":
        bs.nstream(queue, da, db, dc, scalar, n)


def _upload(ab, dev, queue, *arrays):
    bufs = []
    for x in arrays:
        b = ab.alloc_buf(dev, x.dtype, x.size, queue)
        ab.memcpy(queue, b, x)
        bufs.append(b)
    return bufs


def _download(ab, queue, buf, n):
    out = np.empty(n, dtype=buf.dtype)
    ab.memcpy(queue, out, buf, n)
    queue.wait()
    return out


# sizes: empty-ish, ragged (vector tail + scalar tail), one chunk, many chunks, C1's 2^25
SIZES = [1, 3, 31, 1000, 4097, (1 << 16) + 7, (1 << 20), (1 << 22) + 1025, 1 << 25]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", SIZES)
def test_stream_kernels_bit_exact(gpu, dtype, n):
    ab, dev, queue = gpu
    kind = "uniform_f64" if dtype == np.float64 else "uniform_f32"
    a0, b0, c0 = (ol.fill(kind, n, seed=ol.SEED + 10 + k) for k in range(3))
    scalar = 2.0
    for kernel in KERNELS:
        ao, bo, co = a0.copy(), b0.copy(), c0.copy()
        ol.orc_stream(kernel, ao, bo, co, scalar=scalar, init_a=1.0)
        da, db, dc = _upload(ab, dev, queue, a0, b0, c0)
        _run_gpu(ab, queue, kernel, da, db, dc, scalar, 1.0)
        for name, want, buf in (("a", ao, da), ("b", bo, db), ("c", co, dc)):
            got = _download(ab, queue, buf, n)
            assert got.tobytes() == want.tobytes(), f"{kernel}/{name} n={n} differs from the oracle"
        for b in (da, db, dc):
            b.free()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_triad_arbitrary_scalar_no_fma(gpu, dtype):
    """A scalar whose product is inexact: any FMA contraction on either side would show up here."""
    ab, dev, queue = gpu
    n = (1 << 20) + 3
    kind = "uniform_f64" if dtype == np.float64 else "uniform_f32"
    a0, b0, c0 = (ol.fill(kind, n, seed=77 + k) for k in range(3))
    scalar = 0.3
    for kernel in ("mul", "triad", "nstream"):
        ao, bo, co = a0.copy(), b0.copy(), c0.copy()
        ol.orc_stream(kernel, ao, bo, co, scalar=scalar)
        da, db, dc = _upload(ab, dev, queue, a0, b0, c0)
        _run_gpu(ab, queue, kernel, da, db, dc, scalar, 1.0)
        for want, buf in ((ao, da), (bo, db), (co, dc)):
            assert _download(ab, queue, buf, n).tobytes() == want.tobytes(), kernel


def test_reference_sequence_known_answers(gpu):
    """The driver's own check (babelStreamMainTest.cpp:305-368,405): A=1, B=2, C=5, Dot=2N; plus Nstream."""
    ab, dev, queue = gpu
    n = 1 << 22
    for dtype in (np.float64, np.float32):
        da, db, dc = (ab.alloc_buf(dev, dtype, n, queue) for _ in range(3))
        bs = ab.babelstream
        bs.init(queue, da, db, dc)
        bs.copy(queue, da, db)
        bs.mul(queue, da, db)
        bs.add(queue, da, db, dc)
        bs.triad(queue, da, db, dc)
        assert (_download(ab, queue, da, n) == 1).all()
        assert (_download(ab, queue, db, n) == 2).all()
        assert (_download(ab, queue, dc, n) == 5).all()
        assert bs.dot(queue, da, db) == 2 * n
        parts = ab.alloc_buf(dev, dtype, 256, queue)
        bs.dot_partials(queue, da, db, parts)
        assert _download(ab, queue, parts, 256).sum() == 2 * n


def test_misaligned_and_offset_views(gpu):
    """Pointers that are only element-aligned take the scalar path and must still match bit for bit."""
    import ctypes as C

    ab, dev, queue = gpu
    from alpaka_b200 import _lib

    n = 100_003
    a0, b0, c0 = (ol.fill("uniform_f64", n + 1, seed=5 + k) for k in range(3))
    da, db, dc = _upload(ab, dev, queue, a0, b0, c0)
    lib = _lib.load()
    # shift every pointer by one element (8 bytes: not 16-byte aligned)
    _lib.check(lib.b200_stream_triad_f64(queue.handle, da.ptr + 8, db.ptr + 8, dc.ptr + 8, C.c_double(2.0), n))
    want = c0.copy()
    ol.oracle().orc_triad_f64(P(a0[1:]), P(b0[1:]), P(want[1:]), C.c_double(2.0), n)
    got = _download(ab, queue, dc, n + 1)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 5, 1023, (1 << 20) + 17, 1 << 25])
def test_dot_within_tolerance_of_reference_order(gpu, dtype, n):
    ab, dev, queue = gpu
    kind = "uniform_f64" if dtype == np.float64 else "uniform_f32"
    a, b = ol.fill(kind, n, seed=21), ol.fill(kind, n, seed=22)
    sfx = ol.SFX[np.dtype(dtype)]
    # reference order on its CPU back-end: WorkDiv {G,1,1} (G = 8 x 32 blocks) and on its GPU shape {256,1024}
    want_cpu = getattr(ol.oracle(), f"orc_dot_{sfx}")(P(a), P(b), n, 256, 1, None)
    want_gpu_shape = getattr(ol.oracle(), f"orc_dot_{sfx}")(P(a), P(b), n, 256, 1024, None)
    exact = float(np.dot(a.astype(np.float64), b.astype(np.float64))) if n <= (1 << 22) else None
    da, db = _upload(ab, dev, queue, a, b)
    got = float(ab.babelstream.dot(queue, da, db))
    got2 = float(ab.babelstream.dot(queue, da, db))
    assert got == got2, "Dot must be deterministic run to run"
    scale = float(np.sum(np.abs(a.astype(np.float64) * b.astype(np.float64))))
    rtol = DOT_RTOL[np.dtype(dtype)]
    # relative to the magnitude of the terms (the sum itself can cancel to ~0 for U[-1,1) data)
    assert abs(got - float(want_cpu)) <= rtol * scale
    assert abs(got - float(want_gpu_shape)) <= rtol * scale
    if exact is not None:
        assert abs(got - exact) <= rtol * scale
    parts = ab.alloc_buf(dev, dtype, 256, queue)
    ab.babelstream.dot_partials(queue, da, db, parts)
    ph = _download(ab, queue, parts, 256)
    assert abs(float(ph.astype(np.float64).sum()) - float(want_cpu)) <= rtol * scale


@pytest.mark.parametrize("n", [1 << 24, 1 << 25], ids=["2^24", "2^25_C1"])
def test_dot_positive_data_relative_to_value(gpu, n):
    """BASELINE.json north_star, read literally: Dot within 1e-12 RELATIVE TO THE RESULT (positive data, so the sum does not
    cancel), against the reference's order on its CPU back-end ({256,1,1}) -- at C1's 2^25 too (2^30: test_gpu_full_size)."""
    ab, dev, queue = gpu
    a = np.abs(ol.fill("uniform_f64", n, seed=31)) + 0.5
    b = np.abs(ol.fill("uniform_f64", n, seed=32)) + 0.5
    want = ol.oracle().orc_dot_f64(P(a), P(b), n, 256, 1, None)
    da, db = _upload(ab, dev, queue, a, b)
    got = float(ab.babelstream.dot(queue, da, db))
    assert abs(got - want) / abs(want) <= 1e-12


def test_blocking_queue_and_type_errors(gpu):
    ab, dev, _ = gpu
    q = ab.Queue(dev, blocking=True)
    a = ab.alloc_buf(dev, np.float64, 1024, q)
    b = ab.alloc_buf(dev, np.float32, 1024, q)
    with pytest.raises(ab.B200Error):
        ab.babelstream.copy(q, a, b)  # element types differ
    with pytest.raises(ab.B200Error):
        ab.babelstream.copy(q, a, a, n=4096)  # n exceeds the buffer
    c = ab.alloc_buf(dev, np.float64, 1024, q)
    ab.babelstream.init(q, a, c, c, 3.0)
    assert q.empty()  # blocking queue: work is complete when the call returns
    q.close()
