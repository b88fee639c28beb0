"""Sharded GPU paths against the committed golden vectors of the UNMODIFIED reference (tests/golden/multi_gpu_vectors.npz,
made by tests/golden/make_golden_multi.py from oracle/_ref). No oracle call, no /root/reference: usable on the GPU box
from bench.py's N > 1 arm (`parity_n` block), from tests/mp_worker.py and, with every rank on one device, from
tests/test_gpu_golden_multi.py.

What is compared, per world size ("sharded result == unsharded reference result"):
  triad, nstream   every rank's slab window                      == the golden window            bit for bit
  dot (uniform)    fused all-ranks exchange, on EVERY rank       == rank-ordered sum of the per-rank results (bit for bit)
                                                                 and within 1e-12 * sum|a_i b_i| of the reference DotKernel
  dot (integers)   fused exchange on exactly representable data  == the reference DotKernel value bit for bit
  reduce u32/f32   fused exchange (wrap-add / {0,1} data)        == the reference ReduceKernel pair bit for bit
  heat tiles       Py x Px tiles, halo exchange fused into the step kernel, stitched == the undecomposed field
  heat slabs       row slabs, 2 / 3 / 4 / 6 / 8 time levels per launch (>= 2 launches of every depth), stitched == the same field
"""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multi_gpu_vectors.npz")


class Ranks:
    """The ranks of one check. `queues`: rank -> Queue for the ranks that live in THIS process (all of them when one
    process drives every rank on one device; exactly one under torchrun, then `dist` is torch.distributed)."""

    def __init__(self, ab, world: int, queues: dict, dist=None):
        self.ab, self.world, self.queues, self.dist = ab, world, queues, dist
        self.mine = sorted(queues)
        if dist is None and self.mine != list(range(world)):
            raise ValueError("without a process group every rank must live in this process")

    def all_gather(self, per_rank: dict) -> list:
        """One picklable object per local rank -> the list over all ranks, on every process."""
        if self.dist is None:
            return [per_rank[r] for r in range(self.world)]
        got = [None] * self.world
        self.dist.all_gather_object(got, per_rank[self.mine[0]])
        return got

    def connect(self, objs: dict, exchange: bool) -> None:
        from alpaka_b200 import multi

        if self.dist is None:
            (multi.connect_exchange_in_process if exchange else multi.connect_in_process)([objs[r] for r in range(self.world)])
        elif exchange:
            multi.connect_exchange_over_process_group(objs[self.mine[0]], self.dist)
        else:
            multi.connect_over_process_group(objs[self.mine[0]], self.dist)

    def barrier(self) -> None:
        for q in self.queues.values():
            q.wait()
        if self.dist is not None:
            self.dist.barrier()


def _upload(ab, q, host):
    d = ab.alloc_buf(q.dev, host.dtype, host.size, q)
    ab.memcpy(q, d, np.ascontiguousarray(host))
    return d


def _download(ab, q, buf, dtype, n):
    h = np.empty(n, dtype=dtype)
    ab.memcpy(q, h, buf)
    q.wait()
    return h


def check_streams(R: Ranks, G) -> dict:
    from alpaka_b200 import decomp

    ab = R.ab
    a0, b0, c0 = G["stream_in"]
    n = a0.size
    for r in R.mine:
        q = R.queues[r]
        lo, hi = decomp.slab_bounds(n, R.world, r, align=4)
        da, db, dc = (_upload(ab, q, x[lo:hi]) for x in (a0, b0, c0))
        ab.babelstream.triad(q, da, db, dc, 2.0)
        got = _download(ab, q, dc, np.float64, hi - lo)
        if got.tobytes() != G["stream_triad"][lo:hi].tobytes():
            raise AssertionError(f"rank {r}/{R.world}: slab-sharded Triad differs from the reference")
        ab.memcpy(q, dc, np.ascontiguousarray(c0[lo:hi]))
        ab.babelstream.nstream(q, da, db, dc, 2.0)
        got = _download(ab, q, da, np.float64, hi - lo)
        if got.tobytes() != G["stream_nstream"][lo:hi].tobytes():
            raise AssertionError(f"rank {r}/{R.world}: slab-sharded Nstream differs from the reference")
        for d in (da, db, dc):
            d.free()
    return {"triad": "bit-exact", "nstream": "bit-exact"}


def check_dot_reduce(R: Ranks, G) -> dict:
    from alpaka_b200 import decomp, multi

    ab = R.ab
    exs = {r: multi.ScalarExchange(R.queues[r], r, R.world) for r in R.mine}
    R.connect(exs, exchange=True)
    out = {}

    def fused(kind, arrays, dtype):
        """One collective call: every local rank enqueues before anyone waits (the launches wait for each other)."""
        n = arrays[0].size
        bufs, outs = {}, {}
        for r in R.mine:
            q = R.queues[r]
            lo, hi = decomp.slab_bounds(n, R.world, r, align=8)
            bufs[r] = [_upload(ab, q, x[lo:hi]) for x in arrays]
            outs[r] = ab.alloc_buf(q.dev, dtype, 1, q)
        local = {}
        if kind == "dot":
            for r in R.mine:
                local[r] = float(ab.babelstream.dot(R.queues[r], *bufs[r]))
        for r in R.mine:
            if kind == "dot":
                exs[r].dot_async(R.queues[r], bufs[r][0], bufs[r][1], outs[r])
            else:
                exs[r].reduce_sum_async(R.queues[r], bufs[r][0], outs[r])
        got = {r: _download(ab, R.queues[r], outs[r], dtype, 1)[0] for r in R.mine}
        for r in R.mine:
            for d in bufs[r] + [outs[r]]:
                d.free()
        return R.all_gather(got), (R.all_gather(local) if local else None)

    a0, b0, _ = G["stream_in"]
    for rep in range(2):  # twice: the exchange slots are double-buffered by the parity of the call number
        got, local = fused("dot", (a0, b0), np.float64)
        want = decomp.combine_in_rank_order([float(x) for x in local])
        if not all(float(g) == want for g in got):
            raise AssertionError(f"fused Dot exchange {got} != rank-ordered combination {want} (world {R.world})")
        ref = float(G["dot_uniform"][0])
        if abs(want - ref) > 1e-12 * float(np.sum(np.abs(a0 * b0))):
            raise AssertionError(f"sharded Dot {want} outside 1e-12 of the reference DotKernel {ref}")
    out["dot_uniform"] = "== rank-ordered combination (bit-exact on every rank), 1e-12 of the reference"
    ia, ib = G["dot_int_in"]
    got, _ = fused("dot", (ia, ib), np.float64)
    if not all(float(g) == float(G["dot_int"][0]) for g in got):
        raise AssertionError(f"fused Dot on integer data {got} != reference {G['dot_int'][0]}")
    out["dot_int"] = "bit-exact"
    for tag, dtype in (("u32", np.uint32), ("f32", np.float32)):
        got, _ = fused("reduce", (G[f"reduce_{tag}_in"],), dtype)
        if not all(np.array([g], dtype=dtype).tobytes() == G[f"reduce_{tag}"].tobytes() for g in got):
            raise AssertionError(f"fused reduce {tag} exchange {got} != reference {G[f'reduce_{tag}'][0]}")
        out[f"reduce_{tag}"] = "bit-exact"
    R.barrier()
    for r in R.mine:
        if exs[r].status() != 0:
            raise AssertionError(f"rank {r}: a peer's flag never arrived in the fused exchange")
        exs[r].close()
    return out


def check_heat(R: Ranks, G) -> dict:
    from alpaka_b200 import decomp, multi

    ab = R.ab
    want, u0 = G["heat_final"], G["heat_init"]
    NY, NX = want.shape[0] - 2, want.shape[1] - 2
    dx, dy, dt, steps = (float(x) for x in G["heat_params"])
    steps = int(steps)
    out = {}
    corners = np.ones_like(want, dtype=bool)  # never written by either reference kernel (BoundaryKernel.hpp:63-84)
    corners[0, 0] = corners[0, -1] = corners[-1, 0] = corners[-1, -1] = False

    # ---- 2-D tiles, one level per launch
    if len(R.mine) > 1:  # several tiles on one device wait for each other's flags: keep every launch co-resident
        ab.runtime.tune_set("heat.grid_cap", 8)
    try:
        runners = {r: multi.HeatTile(R.queues[r], decomp.tile_for(r, R.world, NY, NX), NY, NX, dt=dt) for r in R.mine}
        R.connect(runners, exchange=False)
        for r in R.mine:
            runners[r].upload(decomp.tile_view(u0, runners[r].tile))
        R.barrier()
        for _ in range(steps):
            for r in R.mine:
                runners[r].step(1)
        R.barrier()
        parts = R.all_gather({r: (runners[r].tile, runners[r].download(), runners[r].status()) for r in R.mine})
        got = np.full_like(want, np.nan)
        for tile, local, status in parts:
            if status != 0:
                raise AssertionError(f"rank {tile.rank}: halo flag wait timed out")
            decomp.stitch(got, tile, local)
        if got[corners].tobytes() != want[corners].tobytes():
            raise AssertionError(f"decomposed heat ({runners[R.mine[0]].tile.py} x {runners[R.mine[0]].tile.px} tiles) differs from the undecomposed reference field")
        t0 = runners[R.mine[0]].tile
        out[f"heat_tiles_{t0.py}x{t0.px}"] = "bit-exact"
        R.barrier()
        for r in R.mine:
            runners[r].close()
    finally:
        if len(R.mine) > 1:
            ab.runtime.tune_set("heat.grid_cap", 0)

    # ---- 2-D tiles with ghost cells 4 deep, four levels per launch (rows inside the launch, columns by the column kernel)
    try:
        deep = {r: multi.HeatTileDeep(R.queues[r], r, R.world, NY, NX, dt=dt, levels=4) for r in R.mine}
    except ab.B200Error:
        deep = None  # tiles smaller than two ghost depths at this world size
    if deep is not None and steps % 4 == 0:
        R.connect(deep, exchange=False)
        for r in R.mine:
            deep[r].upload(deep[r].window(u0))
        R.barrier()
        for _ in range(steps // 4):
            for r in R.mine:
                deep[r].step(4)
        R.barrier()
        parts = R.all_gather({r: (deep[r].tile, deep[r].download(), deep[r].status()) for r in R.mine})
        got = np.full_like(want, np.nan)
        for tile, local, status in parts:
            if status != 0:
                raise AssertionError(f"rank {tile.rank}: deep tile flag wait timed out")
            tile.stitch(got, local)
        if got[corners].tobytes() != want[corners].tobytes():
            raise AssertionError("2-D tiles at four levels per launch differ from the undecomposed reference field")
        t0 = deep[R.mine[0]].tile
        out[f"heat_tiles_{t0.py}x{t0.px}_4_levels"] = "bit-exact"
        R.barrier()
        for r in R.mine:
            deep[r].close()

    # ---- row slabs, 2 / 3 / 4 / 6 / 8 levels per launch and per exchange (4, 6, 8: the walker kernel)
    for levels in (2, 3, 4, 6, 8):
        slabs = {r: multi.HeatSlab(R.queues[r], r, R.world, NY, NX, dt=dt, levels=levels) for r in R.mine}
        R.connect(slabs, exchange=False)
        for r in R.mine:
            slabs[r].upload(slabs[r].window(u0))
        R.barrier()
        for k in decomp.launch_schedule(steps, levels, min_depth=2):
            for r in R.mine:
                slabs[r].step(k)
        R.barrier()
        parts = R.all_gather({r: (slabs[r].tile, slabs[r].download(), slabs[r].status()) for r in R.mine})
        got = np.full_like(want, np.nan)
        for slab, local, status in parts:
            if status != 0:
                raise AssertionError(f"rank {slab.rank}: slab flag wait timed out")
            slab.stitch(got, local)
        if got.tobytes() != want.tobytes():
            raise AssertionError(f"slab-decomposed heat, {levels} levels per launch, differs from the undecomposed reference field")
        out[f"heat_slabs_{levels}_levels"] = "bit-exact"
        R.barrier()
        for r in R.mine:
            slabs[r].close()
    return out


def check_all(R: Ranks) -> dict:
    """Runs every comparison; raises AssertionError on the first mismatch; returns {path: verdict} for the bench line."""
    G = np.load(GOLDEN)
    out = {"world": R.world, "fixture": "tests/golden/multi_gpu_vectors.npz (unmodified reference, oracle/_ref)"}
    out.update(check_streams(R, G))
    out.update(check_dot_reduce(R, G))
    out.update(check_heat(R, G))
    return out
