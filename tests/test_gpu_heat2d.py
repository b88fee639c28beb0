"""Parity of the fused TMA stencil kernel against the oracle's restatement of heatEquation2D: bit-exact
(BASELINE.json asks for <= 1e-12 max-abs after N steps; with contraction pinned and host-side transcendental
tables the difference is exactly zero)."""
import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import P

pytestmark = pytest.mark.gpu


def _run_gpu(ab, queue, u0, steps, dx, dy, dt, fuse=True, **kw):
    ny, nx = u0.shape[0] - 2, u0.shape[1] - 2
    h = ab.heat2d.Heat2D(queue, ny, nx, dx, dy, dt, **kw)
    h.upload(u0)
    h.step(steps, fuse=fuse)
    out = h.download()
    h.close()
    return out


def _init(ny, nx, dx, dy):
    u = np.empty((ny + 2, nx + 2))
    ol.oracle().orc_heat2d_init(P(u), ny, nx, nx + 2, dx, dy)
    return u


# the reference driver's shape (64x64), shapes that are not multiples of the 32x128 tile or of the reference's 16x16
# chunk, single-row / single-column domains, odd widths (last column pair is half valid), and a multi-tile grid
SHAPES = [(64, 64), (1, 1), (1, 7), (9, 1), (16, 16), (33, 129), (31, 127), (100, 257), (256, 1024), (515, 1030)]


@pytest.mark.parametrize("fuse", [1, 2, 3, 4, 6, 8], ids=lambda k: f"{k}_levels_per_launch")
@pytest.mark.parametrize("shape", SHAPES)
def test_heat2d_bit_exact_vs_oracle(gpu, shape, fuse):
    """25 steps: 25 one-step launches; 12 two-level launches (b200_heat2d_step2_f64) + 1; 8 three-level launches
    (b200_heat2d_stepn_f64) + 1; 6 four-level launches + 1; 6 + 6 + 6 + 4 + 3; 8 + 8 + 6 + 3 (4, 6, 8: the walker kernel)."""
    ab, dev, queue = gpu
    ny, nx = shape
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = _init(ny, nx, dx, dy)
    assert ab.heat2d.initial_field(ny, nx, dx, dy).tobytes() == u0.tobytes()
    steps = 25
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    got = _run_gpu(ab, queue, u0, steps, dx, dy, dt, fuse=fuse)
    assert np.max(np.abs(got - want)) <= 1e-12
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("tile", [(32, 8), (32, 16), (32, 32), (64, 16), (64, 32)], ids=lambda t: f"ty{t[0]}_rpt{t[1]}")
def test_heat2d_two_step_kernel_every_tile_shape(gpu, tile):
    """Every instantiation of the two-level kernel (heat.step2_ty / heat.step2_rpt) on a rough field whose extents leave
    partial tiles on both axes: bit-exact against the oracle, ring and corners included."""
    ab, dev, queue = gpu
    ny, nx = 203, 391
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=21).reshape(ny + 2, nx + 2)
    want = ol.orc_heat_run(u0, 1, 6, dx, dy, dt)
    ab.runtime.tune_set("heat.step2_ty", tile[0])
    ab.runtime.tune_set("heat.step2_rpt", tile[1])
    try:
        got = _run_gpu(ab, queue, u0, 6, dx, dy, dt, fuse=True)
    finally:
        ab.runtime.tune_set("heat.step2_ty", 64)
        ab.runtime.tune_set("heat.step2_rpt", 16)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("cfg", [(3, 16, 4), (3, 16, 2), (3, 32, 2), (4, 16, 4), (4, 32, 2)], ids=lambda c: f"levels{c[0]}_rpt{c[1]}_nwy{c[2]}")
def test_heat2d_n_level_kernel_every_tile_shape(gpu, cfg):
    """Every instantiation of the shuffle-exchange N-level kernel on a rough field with partial tiles on both axes."""
    ab, dev, queue = gpu
    levels, rpt, nwy = cfg
    ny, nx = 203, 391
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=22).reshape(ny + 2, nx + 2)
    steps = 2 * levels + 1
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    ab.runtime.tune_set("heat.stepn_rpt", rpt)
    ab.runtime.tune_set("heat.stepn_nwy", nwy)
    ab.runtime.tune_set("heat.walk", 0)  # four levels: the tile kernel, not the walker that replaced it as the default
    try:
        got = _run_gpu(ab, queue, u0, steps, dx, dy, dt, fuse=levels)
    finally:
        ab.runtime.tune_set("heat.stepn_rpt", 16)
        ab.runtime.tune_set("heat.stepn_nwy", 2)
        ab.runtime.tune_set("heat.walk", 1)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("split", [0, 1], ids=["one_kernel", "interior_in_bare_kernel"])
@pytest.mark.parametrize("seg_rows", [0, 40, 7], ids=lambda r: f"seg{r}")
@pytest.mark.parametrize("shape_key", [0, 43, 44], ids=lambda k: f"R{k // 10}_stages{k % 10}")
@pytest.mark.parametrize("levels", [4, 6, 8])
@pytest.mark.parametrize("square", [False, True], ids=["rx_ne_ry", "square_cells"])
def test_heat2d_walker_kernel_every_shape(gpu, square, levels, shape_key, seg_rows, split):
    """Every instantiation of the walker kernel (heatWalkKernel: one warp walks down a 128-column window, levels kept as
    partial sums in registers) on a rough field: partial windows on the right edge, several row segments per window
    (heat.walk_seg_rows forces short ones, down to segments shorter than the 2S-row prologue), chunk counts that do not
    divide the stage ring, both product forms (square cells share v*rX), and with the interior windows' interior rows split
    off into the bare-only kernel launched next to the full one (heat.walk_split = 1, the default shape at 4 and 6 levels).
    Bit-exact against the oracle, ring included."""
    ab, dev, queue = gpu
    ny, nx = (333, 333) if square else (203, 391)
    dx, dy, dt = ol.heat_params(ny, nx)
    assert (dt / (dx * dx) == dt / (dy * dy)) == square
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=24).reshape(ny + 2, nx + 2)
    steps = 2 * levels
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    ab.runtime.tune_set("heat.walk_shape", shape_key)
    ab.runtime.tune_set("heat.walk_seg_rows", seg_rows)
    ab.runtime.tune_set("heat.walk_split", split)
    try:
        got = _run_gpu(ab, queue, u0, steps, dx, dy, dt, fuse=levels)
    finally:
        ab.runtime.tune_set("heat.walk_shape", 0)
        ab.runtime.tune_set("heat.walk_seg_rows", 0)
        ab.runtime.tune_set("heat.walk_split", 0)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("variant", [{}, {"heat.stepn_sq": 0}, {"heat.stepn_ctas": 5}, {"heat.stepn_rpt": 32}, {"heat.stepn_nwy": 4}],
                         ids=lambda v: "default" if not v else "_".join(f"{k.split('.')[1]}{x}" for k, x in v.items()))
@pytest.mark.parametrize("levels", [3, 4])
def test_heat2d_n_level_kernel_square_cells(gpu, levels, variant):
    """Square cells (ny == nx -> rX == rY bit for bit): the N-level kernel shares ONE product v*rX between the horizontal
    and the vertical terms (makeRowN<SQ>). Rough field, partial tiles on both axes; the shared-product kernel, the general
    kernel forced on the same field (heat.stepn_sq = 0) and the five-CTA build must all equal the oracle bit for bit."""
    ab, dev, queue = gpu
    ny = nx = 333
    dx, dy, dt = ol.heat_params(ny, nx)
    assert dt / (dx * dx) == dt / (dy * dy)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=23).reshape(ny + 2, nx + 2)
    steps = 2 * levels + 1
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    for k, v in variant.items():
        ab.runtime.tune_set(k, v)
    ab.runtime.tune_set("heat.walk", 0)  # the tile kernel (the walker has its own test above)
    try:
        got = _run_gpu(ab, queue, u0, steps, dx, dy, dt, fuse=levels)
    finally:
        for k, v in {"heat.stepn_sq": 1, "heat.stepn_ctas": 4, "heat.stepn_rpt": 16, "heat.stepn_nwy": 2, "heat.walk": 1}.items():
            ab.runtime.tune_set(k, v)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("swap", [0, 1], ids=["lds_plain", "lds_swapped_halves"])
@pytest.mark.parametrize("levels,minb", [(4, 4), (4, 3), (6, 3), (6, 2), (8, 3), (8, 2)], ids=lambda v: str(v))
def test_heat2d_walker_kernel_register_budgets(gpu, levels, minb, swap):
    """Both register budgets of every depth (heat.walk_minb: CTAs per SM the allocation is held to) and both shared-memory
    load orders (heat.walk_lds_swap: conflict-free LDS.128 with the halves of a quad swapped on every other group of four
    lanes) on a field wide enough for interior windows (the bare path), misaligned tail windows and several segments."""
    ab, dev, queue = gpu
    ny, nx = 150, 700
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=25).reshape(ny + 2, nx + 2)
    steps = levels + 4
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    ab.runtime.tune_set("heat.walk_minb", minb)
    ab.runtime.tune_set("heat.walk_seg_rows", 48)
    ab.runtime.tune_set("heat.walk_lds_swap", swap)
    try:
        got = _run_gpu(ab, queue, u0, steps, dx, dy, dt, fuse=levels)
    finally:
        ab.runtime.tune_set("heat.walk_minb", 0)
        ab.runtime.tune_set("heat.walk_seg_rows", 0)
        ab.runtime.tune_set("heat.walk_lds_swap", 0)
    assert got.tobytes() == want.tobytes()


def test_heat2d_two_step_refuses_decomposed_tiles(gpu):
    """Ghost sides would need the neighbour's intermediate level: the C ABI refuses instead of computing garbage."""
    ab, dev, queue = gpu
    import ctypes as C

    ny, nx = 40, 72
    dx, dy, dt = ol.heat_params(ny, nx)
    h = ab.heat2d.Heat2D(queue, ny, nx, dx, dy, dt, edges=ab.heat2d.EDGE_TOP | ab.heat2d.EDGE_LEFT)
    rc = ab._lib.load().b200_heat2d_step2_f64(h.plan, queue.handle, 0, h.rx, h.ry, 1.0, 1.0)
    assert rc != 0
    h.close()


def test_heat2d_random_field_bit_exact(gpu):
    """A rough field exercises every neighbour term (the analytic field is smooth)."""
    ab, dev, queue = gpu
    ny, nx = 200, 300
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=4).reshape(ny + 2, nx + 2)
    want = ol.orc_heat_run(u0, 1, 10, dx, dy, dt)
    got = _run_gpu(ab, queue, u0, 10, dx, dy, dt)
    assert got.tobytes() == want.tobytes()


def test_heat2d_reference_known_answer(gpu):
    """heatEquation2D.cpp:54-59 + analyticalSolution.hpp:49: 64x64, 4000 steps, tMax 0.1 -> error < 1e-4."""
    ab, dev, queue = gpu
    ny = nx = 64
    steps, tmax = 4000, 0.1
    dx, dy, dt = 1.0 / (nx + 1), 1.0 / (ny + 1), tmax / steps
    u0 = _init(ny, nx, dx, dy)
    got = _run_gpu(ab, queue, u0, steps, dx, dy, dt)
    err = ab.heat2d.validate_solution(got, dx, dy, tmax)
    assert err < 1e-4
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    assert got.tobytes() == want.tobytes()


def test_heat2d_large_grid_parity_and_properties(gpu):
    """2048^2: a few steps bit-exact against the oracle, then size-independent properties."""
    ab, dev, queue = gpu
    ny = nx = 2048
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = _init(ny, nx, dx, dy)
    want = ol.orc_heat_run(u0, 1, 8, dx, dy, dt)
    got = _run_gpu(ab, queue, u0, 8, dx, dy, dt)
    assert got.tobytes() == want.tobytes()
    # corners never written
    for j, i in ((0, 0), (0, -1), (-1, 0), (-1, -1)):
        assert got[j, i] == u0[j, i]
    # maximum principle for the FTCS scheme under the stability bound: no new extrema in the core
    assert got[1:-1, 1:-1].max() <= u0.max() + 1e-15 and got[1:-1, 1:-1].min() >= min(u0.min(), 0.0) - 1e-15


def test_heat2d_window_split_equals_full_step(gpu):
    """Interior / edge-strip split used for halo overlap: the union of windows equals one full step."""
    ab, dev, queue = gpu
    ny, nx = 130, 300
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=8).reshape(ny + 2, nx + 2)
    want = ol.orc_heat_run(u0, 1, 1, dx, dy, dt)
    h = ab.heat2d.Heat2D(queue, ny, nx, dx, dy, dt)
    h.upload(u0)
    # interior first, then four strips (odd split points on purpose)
    h.step_window(5, ny - 3, 7, nx - 6, advance=False)
    h.step_window(0, 5, 0, nx + 2, advance=False)
    h.step_window(ny - 3, ny + 2, 0, nx + 2, advance=False)
    h.step_window(5, ny - 3, 0, 7, advance=False)
    h.step_window(5, ny - 3, nx - 6, nx + 2, advance=True)
    got = h.download()
    assert got.tobytes() == want.tobytes()


def test_heat2d_ghost_edges_untouched(gpu):
    """Sub-domain form: sides not flagged as physical boundaries are ghost cells and must not be written."""
    ab, dev, queue = gpu
    ny, nx = 40, 72
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=12).reshape(ny + 2, nx + 2)
    full = ol.orc_heat_run(u0, 1, 1, dx, dy, dt)
    got = _run_gpu(ab, queue, u0, 1, dx, dy, dt, edges=ab.heat2d.EDGE_TOP | ab.heat2d.EDGE_LEFT)
    assert got[1:-1, 1:-1].tobytes() == full[1:-1, 1:-1].tobytes()
    assert got[0, 1:-1].tobytes() == full[0, 1:-1].tobytes() and got[1:-1, 0].tobytes() == full[1:-1, 0].tobytes()
    assert got[-1, :].tobytes() == u0[-1, :].tobytes() and got[:, -1].tobytes() == u0[:, -1].tobytes()


def test_heat2d_argument_errors(gpu):
    ab, dev, queue = gpu
    with pytest.raises(ab.B200Error):
        ab.heat2d.Heat2D(queue, 0, 8, 0.1, 0.1, 1e-4)
    with pytest.raises(ab.B200Error):  # stability condition, heatEquation2D.cpp:67-73
        ab.heat2d.Heat2D(queue, 8, 8, 0.1, 0.1, 1.0)


@pytest.mark.parametrize("levels", [4, 8])
def test_heat2d_walker_rows_not_32_byte_aligned(gpu, levels):
    """A row pitch that is a multiple of 16 but not of 32 bytes (the C ABI asks for 16): the walker's 256-bit stores are not
    possible, every quad leaves as two 128-bit stores. Views over one raw allocation, the C ABI called directly."""
    import ctypes as C

    from alpaka_b200 import _lib
    from alpaka_b200.runtime import Buf

    ab, dev, queue = gpu
    ny, nx = 90, 196  # (nx + 2) * 8 = 1584 = 99 * 16 bytes per row
    pitch = (nx + 2) * 8
    assert pitch % 16 == 0 and pitch % 32 != 0
    dx, dy, dt = ol.heat_params(ny, nx)
    u0 = ol.fill("uniform_f64", (ny + 2) * (nx + 2), seed=26).reshape(ny + 2, nx + 2)
    steps = 2 * levels
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    raw = ab.alloc_buf(dev, np.float64, 2 * (ny + 2) * (nx + 2), queue)
    views = [Buf(dev, np.float64, (ny + 2, nx + 2), native_ptr=raw.ptr + b * (ny + 2) * pitch, pitch_bytes=pitch) for b in range(2)]
    for v in views:
        ab.memcpy(queue, v, u0)
    sx, sy = ab.heat2d.boundary_tables(ny, nx, dx, dy)
    lib = _lib.load()
    plan = C.c_void_p()
    _lib.check(lib.b200_heat2d_plan_create(dev.idx, views[0].ptr, views[1].ptr, pitch, ny, nx, sx.ctypes.data, sy.ctypes.data, 15, C.byref(plan)))
    try:
        cur = 0
        for launch in range(2):
            tfs = (C.c_double * levels)(*[ab.heat2d.time_factor(launch * levels + 1 + l, dt) for l in range(levels)])
            _lib.check(lib.b200_heat2d_stepn_f64(plan, queue.handle, cur, dt / (dx * dx), dt / (dy * dy), levels, tfs))
            cur ^= 1
        got = np.empty_like(u0)
        ab.memcpy(queue, got, views[cur])
        queue.wait()
    finally:
        lib.b200_heat2d_plan_destroy(plan)
        raw.free()
    assert got.tobytes() == want.tobytes()
