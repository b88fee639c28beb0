"""Host-side logic of the multi-GPU layer on CPU (gloo, world_size 2 and 4): slab partitioning, the rank-ordered Dot
exchange, and the 2-D heat decomposition (tile geometry, neighbour wiring, which cells travel as halos, stitching).

The per-tile arithmetic here is a numpy restatement of ONE FTCS step in the reference's evaluation order
(StencilKernel.hpp:84-86 + BoundaryKernel.hpp:63-84) -- test code only, used so that the decomposition logic can be
checked without a GPU: the stitched result of the decomposed run must equal the UNDECOMPOSED oracle bit for bit."""
import math
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import oracle_lib as ol
from oracle_lib import P

from alpaka_b200 import decomp


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ------------------------------------------------------------------------------------------------ pure logic
@pytest.mark.parametrize("n,world,align", [(0, 1, 1), (10, 3, 1), (1 << 20, 8, 4), ((1 << 20) + 7, 8, 4), (5, 8, 2), (1000003, 4, 2)])
def test_slab_bounds_partition_exactly(n, world, align):
    seen = 0
    prev_hi = 0
    for r in range(world):
        lo, hi = decomp.slab_bounds(n, world, r, align)
        assert lo == prev_hi and lo <= hi <= n
        assert lo % align == 0 or lo == n
        seen += hi - lo
        prev_hi = hi
    assert seen == n and prev_hi == n


def test_process_grid_shapes():
    assert [decomp.process_grid(w) for w in (1, 2, 4, 8, 6, 3)] == [(1, 1), (2, 1), (2, 2), (4, 2), (3, 2), (3, 1)]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_tiles_cover_the_field_once_and_neighbours_are_mutual(world):
    NY, NX = 64, 96
    owner = np.full((NY + 2, NX + 2), -1)
    tiles = [decomp.tile_for(r, world, NY, NX) for r in range(world)]
    for t in tiles:
        marker = np.full(t.shape, t.rank)
        before = owner.copy()
        decomp.stitch(owner, t, marker)
        assert ((before == -1) | (before == owner)).all(), "two tiles own the same cell"
        for side in decomp.SIDES:
            nb = t.neighbours[side]
            has_edge = bool(t.edges & {"top": 1, "bottom": 2, "left": 4, "right": 8}[side])
            assert (nb is None) == has_edge
            if nb is not None:
                assert tiles[nb].neighbours[decomp.OPPOSITE[side]] == t.rank
                assert tiles[nb].shape == t.shape
    assert (owner >= 0).all()
    with pytest.raises(ValueError):
        decomp.tile_for(0, 4, 63, 96)


def test_rank_ordered_combination_is_a_left_fold():
    parts = [0.1, 0.2, 0.3, 1e16, -1e16]
    assert decomp.combine_in_rank_order(parts) == ((((0.1 + 0.2) + 0.3) + 1e16) + -1e16)


# ------------------------------------------------------------------------------------------------ numpy tile step
def numpy_tile_step(u, tile, step, dx, dy, dt):
    """One FTCS step of a tile in the reference's operation order; ghosts are NOT touched (the exchange fills them)."""
    rX, rY = dt / (dx * dx), dt / (dy * dy)
    k = 1.0 - 2.0 * rX - 2.0 * rY
    c, l, r_, up, dn = u[1:-1, 1:-1], u[1:-1, :-2], u[1:-1, 2:], u[:-2, 1:-1], u[2:, 1:-1]
    out = u.copy()
    out[1:-1, 1:-1] = (((c * k + l * rX) + r_ * rX) + up * rY) + dn * rY
    pi = math.pi
    tf = math.exp(-pi * pi * (step * dt))
    sx = np.array([math.sin(pi * ((i + tile.i_offset) * dx)) for i in range(tile.nx + 2)])
    sy = np.array([math.sin(pi * ((j + tile.j_offset) * dy)) for j in range(tile.ny + 2)])
    if tile.edges & decomp.EDGE_TOP:
        out[0, 1:-1] = tf * (sx[1:-1] + sy[0])
    if tile.edges & decomp.EDGE_BOTTOM:
        out[-1, 1:-1] = tf * (sx[1:-1] + sy[-1])
    if tile.edges & decomp.EDGE_LEFT:
        out[1:-1, 0] = tf * (sx[0] + sy[1:-1])
    if tile.edges & decomp.EDGE_RIGHT:
        out[1:-1, -1] = tf * (sx[-1] + sy[1:-1])
    return out


def oracle_field(NY, NX, steps):
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = np.empty((NY + 2, NX + 2))
    ol.oracle().orc_heat2d_init(P(u0), NY, NX, NX + 2, dx, dy)
    return u0, ol.orc_heat_run(u0, 1, steps, dx, dy, dt), (dx, dy, dt)


def test_numpy_tile_step_is_the_oracle_step_on_an_undecomposed_field():
    NY, NX, steps = 24, 40, 9
    u0, want, (dx, dy, dt) = oracle_field(NY, NX, steps)
    t = decomp.tile_for(0, 1, NY, NX)
    u = u0.copy()
    for s in range(1, steps + 1):
        u = numpy_tile_step(u, t, s, dx, dy, dt)
    assert u.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------------------------ gloo workers
def _worker(rank, world, port, NY, NX, steps, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dx, dy, dt = ol.heat_params(NY, NX)
        u0 = np.empty((NY + 2, NX + 2))
        ol.oracle().orc_heat2d_init(P(u0), NY, NX, NX + 2, dx, dy)
        tile = decomp.tile_for(rank, world, NY, NX)
        u = np.ascontiguousarray(decomp.tile_view(u0, tile)).copy()
        for s in range(1, steps + 1):
            u = numpy_tile_step(u, tile, s, dx, dy, dt)
            # halo exchange: post all sends, then receive (gloo isend/irecv), one message per neighbour
            reqs, recvs = [], []
            for side in decomp.SIDES:
                nb = tile.neighbours[side]
                if nb is None:
                    continue
                send_idx, recv_idx = decomp.halo_slices(tile, side)
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(u[send_idx])), dst=nb, tag=decomp.SIDES.index(side)))
                buf = torch.empty(u[recv_idx].shape, dtype=torch.float64)
                # the neighbour sends from ITS side opposite to ours
                reqs.append(dist.irecv(buf, src=nb, tag=decomp.SIDES.index(decomp.OPPOSITE[side])))
                recvs.append((recv_idx, buf))
            for r in reqs:
                r.wait()
            for idx, buf in recvs:
                u[idx] = buf.numpy()
        # slab-sharded Dot with the rank-ordered exchange
        n = 100003
        lo, hi = decomp.slab_bounds(n, world, rank, align=4)
        a = ol.fill("uniform_f64", hi - lo, seed=1, first=lo)
        b = ol.fill("uniform_f64", hi - lo, seed=2, first=lo)
        part = torch.tensor([float(ol.oracle().orc_dot_f64(P(a), P(b), hi - lo, 64, 1, None))], dtype=torch.float64)
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        total = decomp.combine_in_rank_order([float(p.item()) for p in parts])
        np.save(os.path.join(out_dir, f"tile{rank}.npy"), u)
        np.save(os.path.join(out_dir, f"dot{rank}.npy"), np.array([total]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,NY,NX", [(2, 32, 48), (4, 32, 48)])
def test_decomposed_heat_and_dot_over_gloo(tmp_path, world, NY, NX):
    steps = 15
    port = free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, NY, NX, steps, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "a gloo worker failed"
    _, want, _ = oracle_field(NY, NX, steps)
    out = np.full((NY + 2, NX + 2), np.nan)
    for r in range(world):
        decomp.stitch(out, decomp.tile_for(r, world, NY, NX), np.load(tmp_path / f"tile{r}.npy"))
    mask = np.ones_like(want, dtype=bool)
    mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = False
    assert out[mask].tobytes() == want[mask].tobytes()
    # every rank computed the same Dot, within 1e-12 of the unsharded oracle
    n = 100003
    a, b = ol.fill("uniform_f64", n, seed=1), ol.fill("uniform_f64", n, seed=2)
    d_orc = float(ol.oracle().orc_dot_f64(P(a), P(b), n, 64, 1, None))
    dots = [float(np.load(tmp_path / f"dot{r}.npy")[0]) for r in range(world)]
    assert len(set(dots)) == 1
    assert abs(dots[0] - d_orc) <= 1e-12 * float(np.sum(np.abs(a * b)))


# ------------------------------------------------------------------------------------------------ row slabs, several time levels per exchange
@pytest.mark.parametrize("n,depth,min_depth", [(1000, 3, 2), (10, 3, 2), (13, 4, 2), (12, 2, 2), (7, 3, 1), (5, 1, 1), (4, 3, 1), (0, 3, 2), (2, 4, 2)])
def test_launch_schedule_covers_the_steps(n, depth, min_depth):
    sched = decomp.launch_schedule(n, depth, min_depth)
    assert sum(sched) == n and all(min_depth <= k <= depth for k in sched)
    assert len(sched) <= -(-n // depth) + 1  # at most one launch more than the minimum


def test_launch_schedule_known_answers_and_refusals():
    assert decomp.launch_schedule(1000, 3, 2) == [3] * 332 + [2, 2]
    assert decomp.launch_schedule(4, 3) == [2, 2] and decomp.launch_schedule(7, 4) == [4, 3]
    for bad in ((1, 3, 2), (3, 2, 2)):  # a slab cannot advance a single level
        with pytest.raises(ValueError):
            decomp.launch_schedule(*bad)


@pytest.mark.parametrize("world,ghost", [(1, 2), (2, 2), (2, 3), (4, 3), (8, 4)])
def test_slabs_cover_the_field_once_and_windows_round_trip(world, ghost):
    NY, NX = 8 * ghost * world, 40
    field = np.arange((NY + 2) * (NX + 2), dtype=np.float64).reshape(NY + 2, NX + 2)
    owner = np.full((NY + 2, NX + 2), -1.0)
    out = np.full((NY + 2, NX + 2), np.nan)
    slabs = [decomp.slab_for(r, world, NY, NX, ghost) for r in range(world)]
    for s in slabs:
        w = s.window(field)
        assert w.shape == s.shape == (NY // world + 2 * ghost, NX + 2)
        j0, j1 = s.owned_rows()
        assert (owner[s.g0 + j0 : s.g0 + j1] == -1).all(), "two slabs own the same row"
        s.stitch(owner, np.full(s.shape, float(s.rank)))
        s.stitch(out, w)
        for side, other in (("top", "bottom"), ("bottom", "top")):
            nb = s.neighbours[side]
            assert (nb is None) == bool(s.edges & {"top": 1, "bottom": 2}[side])
            if nb is not None:
                assert slabs[nb].neighbours[other] == s.rank
                # what I send up/down is exactly what the neighbour holds in its ghost rows on the opposite side
                assert w[s.send_rows(side)].tobytes() == slabs[nb].window(field)[slabs[nb].recv_rows(other)].tobytes()
        assert s.neighbours["left"] is None and s.neighbours["right"] is None and s.edges & 4 and s.edges & 8
    assert (owner >= 0).all() and out.tobytes() == field.tobytes()
    with pytest.raises(ValueError):
        decomp.slab_for(0, 3, 64, 8, 2)  # rows do not divide
    with pytest.raises(ValueError):
        decomp.slab_for(0, 2, 8, 8, 3)  # slabs shallower than two ghost depths


def numpy_slab_launch(u, slab, first_step, levels, dx, dy, dt):
    """`levels` FTCS steps of one slab without communication, the way the fused kernels do it: level l is produced on the
    core rows plus the (levels - l) rows beyond them that deeper levels need on a side with a neighbour; ring columns of
    those rows and the physical ring rows take the analytic boundary value of that level. Ghost rows keep level 0."""
    G, ny = slab.ghost, slab.ny
    rX, rY = dt / (dx * dx), dt / (dy * dy)
    k = 1.0 - 2.0 * rX - 2.0 * rY
    pi = math.pi
    sx = np.array([math.sin(pi * (i * dx)) for i in range(slab.nx + 2)])
    sy = np.array([math.sin(pi * ((slab.g0 + j) * dy)) for j in range(ny + 2 * G)])
    g_top, g_bot = (0 if slab.edges & decomp.EDGE_TOP else 1), (0 if slab.edges & decomp.EDGE_BOTTOM else 1)
    cur = u
    for lvl in range(1, levels + 1):
        tf = math.exp(-pi * pi * ((first_step + lvl - 1) * dt))
        lo, hi = G - (levels - lvl) * g_top, ny + G - 1 + (levels - lvl) * g_bot  # rows on which this level is defined
        nxt = cur.copy()
        rows = slice(lo, hi + 1)
        c, l, r_ = cur[rows, 1:-1], cur[rows, :-2], cur[rows, 2:]
        up, dn = cur[lo - 1 : hi, 1:-1], cur[lo + 1 : hi + 2, 1:-1]
        nxt[rows, 1:-1] = (((c * k + l * rX) + r_ * rX) + up * rY) + dn * rY
        nxt[rows, 0] = tf * (sx[0] + sy[rows])
        nxt[rows, -1] = tf * (sx[-1] + sy[rows])
        if not g_top:
            nxt[G - 1, 1:-1] = tf * (sx[1:-1] + sy[G - 1])
        if not g_bot:
            nxt[ny + G, 1:-1] = tf * (sx[1:-1] + sy[ny + G])
        cur = nxt
    out = u.copy()
    j0, j1 = slab.owned_rows()
    out[j0:j1] = cur[j0:j1]
    # corners of the global field are never written (BoundaryKernel.hpp:63-84)
    if not g_top:
        out[G - 1, 0], out[G - 1, -1] = u[G - 1, 0], u[G - 1, -1]
    if not g_bot:
        out[ny + G, 0], out[ny + G, -1] = u[ny + G, 0], u[ny + G, -1]
    return out


def _slab_worker(rank, world, port, NY, NX, steps, ghost, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dx, dy, dt = ol.heat_params(NY, NX)
        u0 = ol.fill("uniform_f64", (NY + 2) * (NX + 2), seed=77).reshape(NY + 2, NX + 2)
        slab = decomp.slab_for(rank, world, NY, NX, ghost)
        u = slab.window(u0)
        done = 0
        for levels in decomp.launch_schedule(steps, ghost, min_depth=2):
            u = numpy_slab_launch(u, slab, done + 1, levels, dx, dy, dt)
            done += levels
            reqs, recvs = [], []
            for side, other in (("top", "bottom"), ("bottom", "top")):
                nb = slab.neighbours[side]
                if nb is None:
                    continue
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(u[slab.send_rows(side)])), dst=nb, tag=0 if side == "top" else 1))
                buf = torch.empty(u[slab.recv_rows(side)].shape, dtype=torch.float64)
                reqs.append(dist.irecv(buf, src=nb, tag=0 if other == "top" else 1))
                recvs.append((slab.recv_rows(side), buf))
            for r in reqs:
                r.wait()
            for rows, buf in recvs:
                u[rows] = buf.numpy()
        np.save(os.path.join(out_dir, f"slab{rank}.npy"), u)
    finally:
        dist.destroy_process_group()


def test_numpy_slab_launch_is_the_oracle_on_an_undecomposed_field():
    NY, NX, steps, ghost = 24, 40, 11, 3
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = ol.fill("uniform_f64", (NY + 2) * (NX + 2), seed=78).reshape(NY + 2, NX + 2)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    slab = decomp.slab_for(0, 1, NY, NX, ghost)
    u, done = slab.window(u0), 0
    for levels in decomp.launch_schedule(steps, ghost, min_depth=2):
        u = numpy_slab_launch(u, slab, done + 1, levels, dx, dy, dt)
        done += levels
    out = np.full_like(want, np.nan)
    slab.stitch(out, u)
    assert out.tobytes() == want.tobytes()


@pytest.mark.parametrize("world,ghost,steps", [(2, 3, 11), (4, 2, 8), (2, 4, 13)])
def test_slab_decomposed_heat_over_gloo(tmp_path, world, ghost, steps):
    """Row slabs with ghost rows `ghost` deep, several time levels per exchange, launches of mixed depth: the stitched
    result equals the undecomposed oracle bit for bit (host-side geometry and exchange logic of multi.HeatSlab)."""
    NY, NX = 16 * world, 40
    port = free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, NY, NX, steps, ghost, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "a gloo worker failed"
    dx, dy, dt = ol.heat_params(NY, NX)
    u0 = ol.fill("uniform_f64", (NY + 2) * (NX + 2), seed=77).reshape(NY + 2, NX + 2)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    out = np.full((NY + 2, NX + 2), np.nan)
    for r in range(world):
        decomp.slab_for(r, world, NY, NX, ghost).stitch(out, np.load(tmp_path / f"slab{r}.npy"))
    assert out.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------------------------ property tests
from hypothesis import given, settings
from hypothesis import strategies as st


@settings(max_examples=300, deadline=None)
@given(n=st.integers(0, 5000), depth=st.integers(1, 4), min_depth=st.integers(1, 2))
def test_launch_schedule_properties(n, depth, min_depth):
    if depth < min_depth:
        depth = min_depth
    impossible = min_depth == 2 and (n == 1 or (depth == 2 and n % 2 == 1))
    if impossible:
        with pytest.raises(ValueError):
            decomp.launch_schedule(n, depth, min_depth)
        return
    sched = decomp.launch_schedule(n, depth, min_depth)
    assert sum(sched) == n and all(min_depth <= k <= depth for k in sched)
    assert len(sched) <= -(-n // depth) + 1
    assert sched == sorted(sched, reverse=True)  # deep launches first, the shallow tail last


@settings(max_examples=200, deadline=None)
@given(world=st.integers(1, 8), ghost=st.integers(1, 4), rows_per_ghost=st.integers(2, 6), NX=st.integers(1, 50), data=st.data())
def test_slab_geometry_properties(world, ghost, rows_per_ghost, NX, data):
    ny = ghost * rows_per_ghost
    NY = ny * world
    rank = data.draw(st.integers(0, world - 1))
    s = decomp.slab_for(rank, world, NY, NX, ghost)
    assert s.shape == (ny + 2 * ghost, NX + 2) and s.g0 == rank * ny - (ghost - 1)
    j0, j1 = s.owned_rows()
    # owned rows in global padded coordinates: the core rows, plus ring row 0 on the first / NY+1 on the last slab
    lo, hi = s.g0 + j0, s.g0 + j1
    assert lo == (0 if rank == 0 else rank * ny + 1) and hi == (NY + 2 if rank == world - 1 else (rank + 1) * ny + 1)
    # ghost rows of neighbouring slabs are exactly the other's border rows
    if rank + 1 < world:
        t = decomp.slab_for(rank + 1, world, NY, NX, ghost)
        mine = range(s.g0 + s.send_rows("bottom").start, s.g0 + s.send_rows("bottom").stop)
        theirs = range(t.g0 + t.recv_rows("top").start, t.g0 + t.recv_rows("top").stop)
        assert list(mine) == list(theirs)
        back = range(t.g0 + t.send_rows("top").start, t.g0 + t.send_rows("top").stop)
        here = range(s.g0 + s.recv_rows("bottom").start, s.g0 + s.recv_rows("bottom").stop)
        assert list(back) == list(here)


@settings(max_examples=200, deadline=None)
@given(py=st.integers(1, 4), px=st.integers(1, 4), ghost=st.sampled_from([4, 6, 8]), ky=st.integers(2, 5), kx=st.integers(2, 5))
def test_deep_tile_geometry_properties(py, px, ghost, ky, kx):
    """decomp.deep_tile_for: the windows of a global field, stitched back from every tile's OWNED cells, give the field
    again exactly once; ghost cells on a neighbour side are the neighbour's border core cells."""
    ny, nx = ghost * ky, ghost * kx
    NY, NX, world = ny * py, nx * px, py * px
    g = np.arange((NY + 2) * (NX + 2), dtype=np.float64).reshape(NY + 2, NX + 2)
    out = np.full_like(g, np.nan)
    count = np.zeros_like(g)
    tiles = [decomp.deep_tile_for(r, world, NY, NX, ghost, (py, px)) for r in range(world)]
    for t in tiles:
        w = t.window(g)
        assert w.shape == (ny + 2 * ghost, nx + 2 * ghost)
        t.stitch(out, w)
        j0, j1, i0, i1 = t.owned()
        count[t.gj0 + j0 : t.gj0 + j1, t.gi0 + i0 : t.gi0 + i1] += 1
        # the top ghost rows of a tile with an upper neighbour are that neighbour's last `ghost` core rows
        up = t.neighbours["top"]
        if up is not None:
            wu = tiles[up].window(g)
            assert np.array_equal(w[0:ghost, ghost:-ghost], wu[ny : ny + ghost, ghost:-ghost])
        left = t.neighbours["left"]
        if left is not None:
            wl = tiles[left].window(g)
            assert np.array_equal(w[:, 0:ghost], wl[:, nx : nx + ghost])
    assert np.array_equal(out, g) and np.all(count == 1)
    with pytest.raises(ValueError):
        decomp.deep_tile_for(0, world, NY, NX, max(ny, nx))  # tiles smaller than two ghost depths
