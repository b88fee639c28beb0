"""Pins the C oracle (oracle/hotpath_oracle.c) to the UNMODIFIED reference compiled into oracle/_ref, and both to
the reference's own known answers. CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import P

N = 1 << 18


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", ["init", "copy", "mul", "add", "triad", "nstream"])
@pytest.mark.parametrize("acc", [0, 1])
def test_stream_kernels_bit_exact(ref, kernel, dtype, acc):
    kind = "uniform_f64" if dtype == np.float64 else "uniform_f32"
    a0, b0, c0 = (ol.fill(kind, N, seed=ol.SEED + k) for k in range(3))
    ao, bo, co = a0.copy(), b0.copy(), c0.copy()
    ar, br, cr = a0.copy(), b0.copy(), c0.copy()
    ol.orc_stream(kernel, ao, bo, co, scalar=2.0, init_a=1.0)
    ol.ref_stream(kernel, ar, br, cr, acc=acc, init_a=1.0)
    for o, r in ((ao, ar), (bo, br), (co, cr)):
        assert o.tobytes() == r.tobytes()


def test_reference_known_answers_babelstream(ref):
    """A=1, B=2, C=5 after Init,Copy,Mult,Add,Triad and Dot == 2N (babelStreamMainTest.cpp:353-355,405)."""
    n = 1 << 17
    for dtype in (np.float64, np.float32):
        a, b, c = (np.empty(n, dtype=dtype) for _ in range(3))
        for k in ("init", "copy", "mul", "add", "triad"):
            ol.orc_stream(k, a, b, c)
        assert (a == 1).all() and (b == 2).all() and (c == 5).all()
        sfx = ol.SFX[np.dtype(dtype)]
        d = getattr(ol.oracle(), f"orc_dot_{sfx}")(P(a), P(b), n, 256, 1024, None)
        assert d == 2 * n
        dr = ref.ref_babelstream_dot(1, 1 if dtype == np.float64 else 0, P(a), P(b), n, 256, None)
        assert dr == 2 * n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("grid", [1, 7, 64, 256])
def test_dot_bit_exact_against_reference_cpu_workdiv(ref, dtype, grid):
    """Reference CPU accs only accept one thread per block: WorkDiv {G,1,1}."""
    kind = "uniform_f64" if dtype == np.float64 else "uniform_f32"
    a, b = ol.fill(kind, N + 13, seed=1), ol.fill(kind, N + 13, seed=2)
    sfx = ol.SFX[np.dtype(dtype)]
    po = np.empty(grid, dtype=dtype)
    pr = np.empty(grid, dtype=dtype)
    do = getattr(ol.oracle(), f"orc_dot_{sfx}")(P(a), P(b), a.size, grid, 1, P(po))
    for acc in (0, 1):
        dr = ref.ref_babelstream_dot(acc, 1 if dtype == np.float64 else 0, P(a), P(b), a.size, grid, P(pr))
        assert po.tobytes() == pr.tobytes()
        assert np.array(do, dtype=dtype).tobytes() == np.array(dr, dtype=dtype).tobytes()


def _ref_block_count(ref, n, acc):
    mp = 1 if acc == 0 else ref.ref_omp_max_threads()
    return ol.oracle().orc_reduce_block_count(n, mp, 1)


@pytest.mark.parametrize("dtype", [np.uint32, np.int32, np.uint64, np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, (1 << 18) + 5])
@pytest.mark.parametrize("acc", [0, 1])
def test_reduce_bit_exact(ref, dtype, n, acc):
    rng = np.random.default_rng(n)
    if np.dtype(dtype).kind == "f":
        x = rng.random(n).astype(dtype)
    else:
        x = rng.integers(0, 2**31 - 1, n).astype(dtype)
    got = ol.orc_reduce(x, _ref_block_count(ref, n, acc), 1, iterator=0)
    want = ol.ref_reduce(x, acc)
    assert np.array(got).tobytes() == np.array(want).tobytes()


def test_reduce_reference_closed_form(ref):
    """reduce.cpp:137-148: x[i] = i+1, expected n/2*(n+1) mod 2^32, on both iterators of the oracle."""
    n = 1 << 20
    x = (np.arange(n, dtype=np.uint64) + 1).astype(np.uint32)
    expected = np.uint32((n // 2 * (n + 1)) % 2**32)
    assert ol.ref_reduce(x, 0) == expected
    assert ol.orc_reduce(x, ol.oracle().orc_reduce_block_count(n, 1, 1), 1, 0) == expected
    assert ol.orc_reduce(x, ol.oracle().orc_reduce_block_count(n, 148, 256), 256, 1) == expected


@pytest.mark.parametrize("shape", [(16, 16), (32, 48), (64, 64)])
@pytest.mark.parametrize("acc", [0, 1])
def test_heat2d_bit_exact(ref, shape, acc):
    ny, nx = shape
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = np.empty((ny + 2, nx + 2))
    ref.ref_heat2d_init(P(u0), ny, nx, dx, dy)
    uo = np.empty_like(u0)
    ol.oracle().orc_heat2d_init(P(uo), ny, nx, nx + 2, dx, dy)
    assert uo.tobytes() == u0.tobytes()
    steps = 50
    ur = u0.copy()
    assert ref.ref_heat2d_run(acc, P(ur), ny, nx, 1, steps, dx, dy, dt, None) == 0
    uo = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    assert uo.tobytes() == ur.tobytes()
    e_o = ol.oracle().orc_heat2d_validate(P(uo), ny, nx, nx + 2, dx, dy, steps * dt)
    e_r = ref.ref_heat2d_validate(P(ur), ny, nx, dx, dy, steps * dt)
    assert e_o == e_r


def test_heat2d_reference_known_answer(ref):
    """The shipped driver's configuration: 64x64, 4000 steps, tMax 0.1 -> max-abs error < 1e-4
    (heatEquation2D.cpp:54-59, analyticalSolution.hpp:49)."""
    ny = nx = 64
    steps, tmax = 4000, 0.1
    dx, dy, dt = 1.0 / (nx + 1), 1.0 / (ny + 1), tmax / steps
    u = np.empty((ny + 2, nx + 2))
    ol.oracle().orc_heat2d_init(P(u), ny, nx, nx + 2, dx, dy)
    u = ol.orc_heat_run(u, 1, steps, dx, dy, dt)
    err = ol.oracle().orc_heat2d_validate(P(u), ny, nx, nx + 2, dx, dy, tmax)
    assert err < 1e-4


def test_separable_boundary_tables_equal_exact_solution(ref):
    """The factorisation fed to the CUDA kernel: tf*(sx[i]+sy[j]) == exactSolution(i*dx, j*dy, step*dt) bit for bit."""
    ny, nx = 40, 24
    dx, dy, dt = ol.heat_params(ny, nx)
    sx, sy = np.empty(nx + 2), np.empty(ny + 2)
    ol.oracle().orc_heat2d_boundary_tables(P(sx), P(sy), ny, nx, dx, dy)
    for step in (0, 1, 7, 1000):
        tf = ol.oracle().orc_heat2d_time_factor(step, dt)
        for j in (0, 1, ny, ny + 1):
            for i in (0, 1, nx, nx + 1):
                want = ref.ref_heat2d_exact(i * dx, j * dy, step * dt)
                assert tf * (sx[i] + sy[j]) == want
