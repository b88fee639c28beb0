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
