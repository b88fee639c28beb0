"""Golden vectors generated from the UNMODIFIED reference (tests/golden/make_golden.py -> reference_vectors.npz):
CPU tests pin the C oracle to them, GPU tests pin the CUDA path (through the C ABI) to them. Bit-exact everywhere
except Dot, whose summation order legitimately differs on the GPU (tolerance 1e-12 / 1e-5 relative, BASELINE.json)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
KERNELS = ["init", "copy", "mul", "add", "triad", "nstream"]
DT = {"f64": np.float64, "f32": np.float32, "u32": np.uint32, "i32": np.int32, "u64": np.uint64}
REDUCE_CASES = [(t, n) for t in ("u32", "i32", "u64", "f32", "f64") for n in (1, 17, 1000, 20011)]
HEAT_CASES = [(16, 16, 100), (32, 48, 25)]


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("kernel", KERNELS)
def test_oracle_stream_matches_reference_vectors(kernel, tag):
    a, b, c = (x.copy() for x in G[f"stream_{tag}_in"])
    ol.orc_stream(kernel, a, b, c, scalar=2.0, init_a=1.0)
    assert np.stack([a, b, c]).tobytes() == G[f"stream_{tag}_{kernel}"].tobytes()


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("grid", [1, 7, 256])
def test_oracle_dot_matches_reference_vectors(grid, tag):
    a, b, _ = G[f"stream_{tag}_in"]
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    partials = np.empty(grid, dtype=DT[tag])
    d = getattr(ol.oracle(), f"orc_dot_{tag}")(P(a), P(b), a.size, grid, 1, P(partials))
    assert partials.tobytes() == G[f"dot_{tag}_g{grid}_partials"].tobytes()
    assert np.array([d], dtype=DT[tag]).tobytes() == G[f"dot_{tag}_g{grid}"].tobytes()


@pytest.mark.parametrize("tag,n", REDUCE_CASES)
def test_oracle_reduce_matches_reference_vectors(tag, n):
    x = np.ascontiguousarray(G[f"reduce_{tag}_n{n}_in"])
    got = ol.orc_reduce(x, ol.oracle().orc_reduce_block_count(n, 1, 1), 1, iterator=0)
    assert np.array([got], dtype=DT[tag]).tobytes() == G[f"reduce_{tag}_n{n}"].tobytes()


@pytest.mark.parametrize("ny,nx,steps", HEAT_CASES)
def test_oracle_heat_matches_reference_vectors(ny, nx, steps):
    dx, dy, dt, s = G[f"heat_{ny}x{nx}_params"]
    assert int(s) == steps
    u0 = np.empty((ny + 2, nx + 2))
    ol.oracle().orc_heat2d_init(P(u0), ny, nx, nx + 2, dx, dy)
    assert u0.tobytes() == G[f"heat_{ny}x{nx}_init"].tobytes()
    got = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    assert got.tobytes() == G[f"heat_{ny}x{nx}_s{steps}"].tobytes()


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("kernel", KERNELS)
def test_gpu_stream_matches_reference_vectors(gpu, kernel, tag):
    ab, dev, q = gpu
    bs = ab.babelstream
    host = [np.ascontiguousarray(x) for x in G[f"stream_{tag}_in"]]
    n = host[0].size
    bufs = [ab.alloc_buf(dev, DT[tag], n, q) for _ in range(3)]
    for d, h in zip(bufs, host):
        ab.memcpy(q, d, h)
    a, b, c = bufs
    {"init": lambda: bs.init(q, a, b, c, 1.0), "copy": lambda: bs.copy(q, a, b), "mul": lambda: bs.mul(q, a, b, 2.0),
     "add": lambda: bs.add(q, a, b, c), "triad": lambda: bs.triad(q, a, b, c, 2.0),
     "nstream": lambda: bs.nstream(q, a, b, c, 2.0)}[kernel]()
    got = [np.empty(n, dtype=DT[tag]) for _ in range(3)]
    for h, d in zip(got, bufs):
        ab.memcpy(q, h, d)
    q.wait()
    assert np.stack(got).tobytes() == G[f"stream_{tag}_{kernel}"].tobytes()
    for d in bufs:
        d.free()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,rtol", [("f64", 1e-12), ("f32", 1e-5)])
def test_gpu_dot_matches_reference_vectors(gpu, tag, rtol):
    ab, dev, q = gpu
    a, b, _ = (np.ascontiguousarray(x) for x in G[f"stream_{tag}_in"])
    da, db = ab.alloc_buf(dev, DT[tag], a.size, q), ab.alloc_buf(dev, DT[tag], a.size, q)
    ab.memcpy(q, da, a)
    ab.memcpy(q, db, b)
    got = float(ab.babelstream.dot(q, da, db))
    scale = float(np.sum(np.abs(a.astype(np.float64) * b.astype(np.float64))))
    for grid in (1, 7, 256):
        assert abs(got - float(G[f"dot_{tag}_g{grid}"][0])) <= rtol * scale
    da.free()
    db.free()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n", REDUCE_CASES)
def test_gpu_reduce_matches_reference_vectors(gpu, tag, n):
    ab, dev, q = gpu
    x = np.ascontiguousarray(G[f"reduce_{tag}_n{n}_in"])
    d = ab.alloc_buf(dev, DT[tag], n, q)
    ab.memcpy(q, d, x)
    got = ab.reduce.reduce_sum(q, d)
    # integer sums wrap (order-free); the float inputs are {0,1} so every order gives the same exact sum
    assert np.array([got], dtype=DT[tag]).tobytes() == G[f"reduce_{tag}_n{n}"].tobytes()
    d.free()


@pytest.mark.gpu
@pytest.mark.parametrize("ny,nx,steps", HEAT_CASES)
def test_gpu_heat_matches_reference_vectors(gpu, ny, nx, steps):
    ab, dev, q = gpu
    dx, dy, dt, _ = G[f"heat_{ny}x{nx}_params"]
    h = ab.heat2d.Heat2D(q, ny, nx, float(dx), float(dy), float(dt))
    h.upload(G[f"heat_{ny}x{nx}_init"])
    h.step(steps)
    got = h.download()
    h.close()
    assert got.tobytes() == G[f"heat_{ny}x{nx}_s{steps}"].tobytes()
