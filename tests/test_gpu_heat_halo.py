"""2-D decomposed heatEquation2D with the halo exchange fused into the step kernel (b200_heat2d_step_halo_f64).

Parity definition ("multi-GPU sharding: parity unpinned in the reference"): the stitched result of
the decomposed run equals the UNDECOMPOSED oracle run bit for bit.

These tests put several tiles on ONE device (one queue per tile, plain device pointers as "peer" pointers): the kernels
of different tiles wait for each other's flag words, so they must be co-resident -- heat.grid_cap bounds every launch
to a few CTAs. The multi-process / multi-device form of the same protocol (CUDA IPC) is test_gpu_multi_process.py."""
import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import P

pytestmark = pytest.mark.gpu


def run_decomposed(ab, dev, NY, NX, grid, steps, cap=8):
    from alpaka_b200 import decomp, multi

    world = grid[0] * grid[1]
    ab.runtime.tune_set("heat.grid_cap", cap)
    try:
        queues = [ab.Queue(dev) for _ in range(world)]
        tiles = [decomp.tile_for(r, world, NY, NX, grid) for r in range(world)]
        runners = [multi.HeatTile(q, t, NY, NX) for q, t in zip(queues, tiles)]
        multi.connect_in_process(runners)
        for r in runners:
            r.upload(r.initial_field())
        for _ in range(steps):
            for r in runners:
                r.step(1)
        for q in queues:
            q.wait()
        out = np.full((NY + 2, NX + 2), np.nan)
        for r in runners:
            assert r.status() == 0, "a flag wait timed out"
            decomp.stitch(out, r.tile, r.download())
        for r in runners:
            r.close()
        return out
    finally:
        ab.runtime.tune_set("heat.grid_cap", 0)


def oracle_run(NY, NX, steps):
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = np.empty((NY + 2, NX + 2))
    ol.oracle().orc_heat2d_init(P(u0), NY, NX, NX + 2, dx, dy)
    return ol.orc_heat_run(u0, 1, steps, dx, dy, dt)


@pytest.mark.parametrize("grid", [(2, 1), (1, 2), (2, 2), (4, 2)])
def test_decomposed_equals_undecomposed_small_tiles(gpu, grid):
    """Tiles smaller than 2 x (32 x 128): every tile is an edge strip."""
    ab, dev, _ = gpu
    NY, NX, steps = 64 * grid[0], 96 * grid[1], 12
    got = run_decomposed(ab, dev, NY, NX, grid, steps)
    want = oracle_run(NY, NX, steps)
    # corners of the global field are never written by anyone (as in the reference); compare everything else
    mask = np.ones_like(want, dtype=bool)
    mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = False
    assert got[mask].tobytes() == want[mask].tobytes()


@pytest.mark.parametrize("grid", [(2, 1), (2, 2)])
def test_decomposed_equals_undecomposed_strips_and_interior(gpu, grid):
    """Tiles large enough for the five-window launch: four strips (with peer stores) + interior."""
    ab, dev, _ = gpu
    NY, NX, steps = 160 * grid[0], 512 * grid[1], 9
    got = run_decomposed(ab, dev, NY, NX, grid, steps)
    want = oracle_run(NY, NX, steps)
    mask = np.ones_like(want, dtype=bool)
    mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = False
    assert got[mask].tobytes() == want[mask].tobytes()


def test_single_tile_halo_launch_equals_plain_step(gpu):
    """1 x 1 decomposition: no neighbours; the five-window launch must equal the ordinary fused step."""
    ab, dev, _ = gpu
    NY, NX, steps = 256, 640, 7
    got = run_decomposed(ab, dev, NY, NX, (1, 1), steps, cap=0)
    want = oracle_run(NY, NX, steps)
    assert got.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------------------------------------------
# Row slabs advanced 2, 3, 4, 6 or 8 time levels per launch and per exchange (b200_heat2d_step2_halo_f64 /
# b200_heat2d_stepn_halo_f64): ghost rows as deep as the launch advances.
def run_slabs(ab, dev, NY, NX, world, steps, u0, levels):
    from alpaka_b200 import multi

    queues = [ab.Queue(dev) for _ in range(world)]
    runners = [multi.HeatSlab(q, r, world, NY, NX, levels=levels) for r, q in enumerate(queues)]
    multi.connect_in_process(runners)
    for r in runners:
        r.upload(r.window(u0))
    # launch by launch on every slab in turn (they wait for each other's flags); `sched` = steps per round
    from alpaka_b200 import decomp

    sched = decomp.launch_schedule(steps, levels, min_depth=2)
    assert sum(sched) == steps and all(2 <= k <= levels for k in sched)
    for k in sched:
        for r in runners:
            r.step(k)
    for q in queues:
        q.wait()
    out = np.full((NY + 2, NX + 2), np.nan)
    for r in runners:
        assert r.status() == 0, "a flag wait timed out"
        r.stitch(out, r.download())
    for r in runners:
        r.close()
    return out


# (world, rows per slab, NX): slabs smaller than one 64-row tile (everything is a strip), a slab whose last two core
# rows straddle two tile rows (ny = 127), slabs with interior tile rows and partial tiles in x, a single slab
SLAB_CASES = [(2, 20, 96), (3, 61, 200), (2, 127, 391), (4, 200, 700), (1, 150, 300), (8, 64, 256),
              (2, 128, 256), (4, 75, 300)]  # the last two: square cells (NY == NX), the shared-product kernel


@pytest.mark.parametrize("levels", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("case", SLAB_CASES, ids=lambda c: f"{c[0]}slabs_of_{c[1]}x{c[2]}")
def test_slabs_fused_levels_equal_undecomposed(gpu, case, levels):
    ab, dev, _ = gpu
    world, ny, NX = case
    NY, steps = ny * world, (12 if levels <= 4 else 2 * levels + 4)
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = np.empty((NY + 2, NX + 2))
    ol.oracle().orc_heat2d_init(P(u0), NY, NX, NX + 2, dx, dy)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    got = run_slabs(ab, dev, NY, NX, world, steps, u0, levels)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("levels", [2, 3, 4, 6, 8])
def test_slabs_rough_field_bit_exact(gpu, levels):
    """A rough field exercises every neighbour term across the slab borders (the analytic field is smooth)."""
    ab, dev, _ = gpu
    world, ny, NX, steps = 3, 70, 263, (12 if levels <= 4 else 2 * levels + 2)
    NY = ny * world
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = ol.fill("uniform_f64", (NY + 2) * (NX + 2), seed=33).reshape(NY + 2, NX + 2)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    got = run_slabs(ab, dev, NY, NX, world, steps, u0, levels)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("levels,steps", [(3, 10), (3, 11), (4, 13), (4, 10), (6, 17), (8, 21), (8, 13)])
def test_slabs_mixed_depth_launches(gpu, levels, steps):
    """A step count that is not a multiple of the ghost depth: shallower launches finish it (3+3+2+2, 3+3+3+2, 6+6+3+2,
    8+8+3+2, 8+3+2 ...; walker and tile kernels mixed on one pair of buffers and one flag protocol)."""
    ab, dev, _ = gpu
    world, ny, NX = 3, 70, 263
    NY = ny * world
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = ol.fill("uniform_f64", (NY + 2) * (NX + 2), seed=35).reshape(NY + 2, NX + 2)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    got = run_slabs(ab, dev, NY, NX, world, steps, u0, levels)
    assert got.tobytes() == want.tobytes()


def test_slab_argument_errors(gpu):
    ab, dev, queue = gpu
    from alpaka_b200 import multi

    with pytest.raises(ab.B200Error):  # a slab cannot advance a single level
        s = multi.HeatSlab(queue, 0, 1, 64, 64, levels=3)
        multi.connect_in_process([s])
        try:
            s.step(1)
        finally:
            s.close()
    with pytest.raises(ab.B200Error):  # rows do not divide
        multi.HeatSlab(queue, 0, 3, 64, 64)


# ------------------------------------------------------------------------------------------------------------------
# 2-D tiles advanced 4, 6 or 8 time levels per launch (b200_heat2d_tile_plan_create / b200_heat2d_stepn_tile_f64): ghost
# cells that deep on all four sides, rows exchanged inside the walker launch, columns (and through them the corners) by the
# column kernel that follows it.
def run_deep_tiles(ab, dev, NY, NX, grid, steps, u0, levels):
    from alpaka_b200 import multi

    world = grid[0] * grid[1]
    queues = [ab.Queue(dev) for _ in range(world)]
    runners = [multi.HeatTileDeep(q, r, world, NY, NX, levels=levels, grid=grid) for r, q in enumerate(queues)]
    multi.connect_in_process(runners)
    for r in runners:
        r.upload(r.window(u0))
    left = steps
    while left > 0:  # launch by launch on every tile in turn (they wait for each other's flags)
        k = levels if left >= levels else left
        for r in runners:
            r.step(k)
        left -= k
    for q in queues:
        q.wait()
    out = np.full((NY + 2, NX + 2), np.nan)
    for r in runners:
        assert r.status() == 0, "a flag wait timed out"
        r.stitch(out, r.download())
    for r in runners:
        r.close()
    return out


# (Py, Px), tile rows, tile columns: tiles narrower than one 128-column window, tiles with interior windows, one column of
# tiles (no column exchange), one row of tiles (no row exchange), the 4 x 2 grid of eight GPUs
DEEP_TILE_CASES = [((2, 2), 40, 48), ((2, 2), 70, 300), ((1, 2), 64, 160), ((2, 1), 64, 160), ((4, 2), 32, 200), ((2, 4), 50, 130),
                   ((3, 3), 33, 141), ((1, 1), 64, 64)]


@pytest.mark.parametrize("rough", [False, True], ids=["analytic_field", "rough_field"])
@pytest.mark.parametrize("levels", [4, 6, 8])
@pytest.mark.parametrize("case", DEEP_TILE_CASES, ids=lambda c: f"{c[0][0]}x{c[0][1]}_tiles_of_{c[1]}x{c[2]}")
def test_deep_tiles_equal_undecomposed(gpu, case, levels, rough):
    ab, dev, _ = gpu
    grid, ny, nx = case
    NY, NX = ny * grid[0], nx * grid[1]
    steps = 3 * levels
    dx, dy, dt = ol.heat_params(NY, NX)
    if rough:  # exercises every neighbour term across the tile borders and the corners (the analytic field is smooth)
        u0 = ol.fill("uniform_f64", (NY + 2) * (NX + 2), seed=41).reshape(NY + 2, NX + 2)
    else:
        u0 = np.empty((NY + 2, NX + 2))
        ol.oracle().orc_heat2d_init(P(u0), NY, NX, NX + 2, dx, dy)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    got = run_deep_tiles(ab, dev, NY, NX, grid, steps, u0, levels)
    mask = np.ones_like(want, dtype=bool)  # corners are written by neither reference kernel (BoundaryKernel.hpp:63-84)
    mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = False
    assert got[mask].tobytes() == want[mask].tobytes()


def test_deep_tiles_mixed_depths_and_refusals(gpu):
    """Ghost cells 8 deep: launches of 8, 6 and 4 levels can follow each other (18 = 8 + 6 + 4); step counts no combination
    of walker depths covers are refused, and so are tiles smaller than two ghost depths."""
    ab, dev, queue = gpu
    from alpaka_b200 import multi

    grid, ny, nx = (2, 2), 48, 150
    NY, NX = ny * grid[0], nx * grid[1]
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = ol.fill("uniform_f64", (NY + 2) * (NX + 2), seed=43).reshape(NY + 2, NX + 2)
    want = ol.orc_heat_run(u0, 1, 18, dx, dy, dt)
    queues = [ab.Queue(dev) for _ in range(4)]
    runners = [multi.HeatTileDeep(q, r, 4, NY, NX, levels=8, grid=grid) for r, q in enumerate(queues)]
    multi.connect_in_process(runners)
    for r in runners:
        r.upload(r.window(u0))
    for k in (8, 6, 4):
        for r in runners:
            r.step(k)
    out = np.full((NY + 2, NX + 2), np.nan)
    for r in runners:
        r.stitch(out, r.download())
    with pytest.raises(ab.B200Error):
        runners[0].step(7)
    for r in runners:
        r.close()
    mask = np.ones_like(want, dtype=bool)
    mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = False
    assert out[mask].tobytes() == want[mask].tobytes()
    with pytest.raises(ab.B200Error):
        multi.HeatTileDeep(queue, 0, 4, 24, 24, levels=8, grid=(2, 2))
