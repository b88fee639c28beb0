"""Parity at BASELINE.json's FULL sizes (configs[1..3]) through size-independent properties plus bit-exact windows, as the
oracle cannot restate 26 GB of streams in seconds:

  C2  BabelStream double, 2^30 elements per array: the reference driver's sequence on constant data (exact sums by the
      device reduction: every partial sum is an integer below 2^53, so the order cannot matter), and a ragged window of
      hashed data near the END of the arrays compared bit for bit with the oracle (64-bit index arithmetic, tail
      handling at 8 GB offsets).
  C3  example/reduce: 2^32 uint32 iota -> n/2*(n+1) mod 2^32 = 2^31 (reduce.cpp:148; the reference's own CPU iterator
      cannot do this size) and 2^30 Bernoulli floats (exact count).
  C4  heatEquation2D double 16384 x 16384: 100 steps bit-exact against the oracle at the full size (fused 3-level launches
      + a 2+2 tail), then on to 1000 steps: max-abs error against the analytic solution, untouched corners."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def test_c2_dot_2pow30_within_1e12_of_the_result(gpu):
    """BASELINE.json north_star, read literally, at C2's size: Dot of 2^30 POSITIVE hashed doubles within 1e-12 relative to the
    result itself, against the reference's summation orders (CPU back-end {256,1,1} and GPU shape {256,1024})."""
    ab, dev, queue = gpu
    from oracle_lib import P

    n = 1 << 30
    a = ol.fill("uniform_f64", n, seed=41)
    np.abs(a, out=a)
    a += 0.5
    b = ol.fill("uniform_f64", n, seed=42)
    np.abs(b, out=b)
    b += 0.5
    want_cpu = float(ol.oracle().orc_dot_f64(P(a), P(b), n, 256, 1, None))
    want_gpu_shape = float(ol.oracle().orc_dot_f64(P(a), P(b), n, 256, 1024, None))
    da, db = (ab.alloc_buf(dev, np.float64, n, queue) for _ in range(2))
    try:
        ab.memcpy(queue, da, a)
        ab.memcpy(queue, db, b)
        got = float(ab.babelstream.dot(queue, da, db))
        assert abs(got - want_cpu) <= 1e-12 * abs(want_cpu), (got, want_cpu)
        assert abs(got - want_gpu_shape) <= 1e-12 * abs(want_gpu_shape), (got, want_gpu_shape)
    finally:
        da.free()
        db.free()
        queue.wait()


def test_c2_babelstream_2pow30_known_answers_and_window_parity(gpu):
    ab, dev, queue = gpu
    bs = ab.babelstream
    n = 1 << 30
    a, b, c = (ab.alloc_buf(dev, np.float64, n, queue) for _ in range(3))
    try:
        # the driver's sequence (babelStreamMainTest.cpp:353-355): A = 1, B = 2, C = 5 after copy, mul, add, triad
        bs.init(queue, a, b, c)
        bs.copy(queue, a, c)
        bs.mul(queue, c, b)  # b = 2 * c = 2
        bs.add(queue, a, b, c)  # c = 3
        bs.triad(queue, b, c, a)  # a = b + 2 c = 8
        assert float(ab.reduce.reduce_sum(queue, a)) == 8.0 * n
        assert float(ab.reduce.reduce_sum(queue, b)) == 2.0 * n
        assert float(ab.reduce.reduce_sum(queue, c)) == 3.0 * n
        assert float(bs.dot(queue, b, c)) == 6.0 * n  # Dot = sum b*c, exact for the same reason
        # a ragged window of hashed data ending 3 elements before the end of the arrays
        w = (1 << 20) + 5
        off = n - w - 3
        ha, hb, hc = (ol.fill("uniform_f64", w, seed=300 + k) for k in range(3))
        views = [ab.create_view(dev, buf.ptr + 8 * off, np.float64, w) for buf in (a, b, c)]
        for v, h in zip(views, (ha, hb, hc)):
            ab.memcpy(queue, v, h)
        bs.triad(queue, a, b, c, 0.3)  # the FULL arrays; only the window is compared element by element
        bs.nstream(queue, a, b, c, 0.3)
        got_c, got_a = np.empty(w), np.empty(w)
        ab.memcpy(queue, got_c, views[2])
        ab.memcpy(queue, got_a, views[0])
        queue.wait()
        want_c = hc.copy()
        ol.orc_stream("triad", ha, hb, want_c, scalar=0.3)
        want_a = ha.copy()
        ol.orc_stream("nstream", want_a, hb, want_c, scalar=0.3)
        assert got_c.tobytes() == want_c.tobytes(), "Triad window at the end of 2^30 elements differs from the oracle"
        assert got_a.tobytes() == want_a.tobytes(), "Nstream window at the end of 2^30 elements differs from the oracle"
    finally:
        for buf in (a, b, c):
            buf.free()
        queue.wait()


def test_c3_reduce_2pow32_uint32_closed_form_and_2pow30_float(gpu):
    ab, dev, queue = gpu
    n = 1 << 32
    x = ab.alloc_buf(dev, np.uint32, n, queue)
    try:
        chunk = 1 << 28
        for k in range(n // chunk):  # x[i] = i + 1 (mod 2^32), reduce.cpp:137-138, uploaded 1 GB at a time
            part = (np.arange(k * chunk + 1, (k + 1) * chunk + 1, dtype=np.uint64) & 0xFFFFFFFF).astype(np.uint32)
            ab.memcpy(queue, ab.create_view(dev, x.ptr + 4 * k * chunk, np.uint32, chunk), part)
            queue.wait()
        assert int(ab.reduce.reduce_sum(queue, x)) == (n // 2 * (n + 1)) % 2**32 == 1 << 31
        # prefix: the reference driver's own size
        m = 1 << 28
        assert int(ab.reduce.reduce_sum(queue, x, n=m)) == (m // 2 * (m + 1)) % 2**32
    finally:
        x.free()
        queue.wait()
    nf = 1 << 30
    xf = ab.alloc_buf(dev, np.float32, nf, queue)
    try:
        chunk = 1 << 27
        ones = 0
        for k in range(nf // chunk):
            part = ol.fill("bernoulli_f32", chunk, seed=11, first=k * chunk)
            ones += int(part.sum(dtype=np.float64))
            ab.memcpy(queue, ab.create_view(dev, xf.ptr + 4 * k * chunk, np.float32, chunk), part)
            queue.wait()
        got = float(ab.reduce.reduce_sum(queue, xf))
        assert abs(got - ones) <= 1e-5 * ones  # BASELINE.json: float reduce within 1e-5 relative
    finally:
        xf.free()
        queue.wait()


def test_c4_heat_16384_bit_exact_100_steps_then_1000_steps_vs_analytic(gpu):
    ab, dev, queue = gpu
    ny = nx = 16384
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ab.heat2d.initial_field(ny, nx, dx, dy)
    h = ab.heat2d.Heat2D(queue, ny, nx, dx, dy, dt)
    try:
        h.upload(u0)
        h.step(100)  # 32 three-level launches + 2 + 2
        got = h.download()
        want = ol.orc_heat_run(u0, 1, 100, dx, dy, dt)
        assert got.tobytes() == want.tobytes(), "16384^2 field after 100 steps differs from the oracle"
        del want
        h.step(900)
        got = h.download()
        assert h.step_index == 1000
        err = ab.heat2d.validate_solution(got, dx, dy, 1000 * dt)
        assert err < 1e-4  # analyticalSolution.hpp:49
        for j, i in ((0, 0), (0, -1), (-1, 0), (-1, -1)):
            assert got[j, i] == u0[j, i]  # corners are never written (BoundaryKernel.hpp:63-84)
        assert got[1:-1, 1:-1].max() <= u0.max() + 1e-15 and got[1:-1, 1:-1].min() >= -1e-15  # maximum principle
    finally:
        h.close()
        queue.wait()
