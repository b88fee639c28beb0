"""Worker of tests/test_gpu_multi_process.py: one process per GPU under torchrun (NCCL). Checks the multi-GPU paths
against the UNSHARDED oracle on the same seeded inputs:
  * slab-sharded Triad (bit-exact) and Dot with the rank-ordered scalar exchange (1e-12),
  * slab-sharded uint32 reduce with wrap-add combination (bit-exact),
  * 2-D decomposed heatEquation2D with the halo exchange fused into the step kernel over CUDA-IPC peer pointers
    (bit-exact after stitching).
Prints one line "MP_WORKER_OK <world>" from rank 0 on success; any mismatch raises."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import alpaka_b200 as ab
    import oracle_lib as ol
    from alpaka_b200 import decomp, multi
    from oracle_lib import P

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = ab.Platform().get_dev_by_idx(local_rank)
    q = ab.Queue(dev)

    # ---- slab-sharded streams + Dot
    n = (1 << 22) + 40
    lo, hi = decomp.slab_bounds(n, world, rank, align=4)
    m = hi - lo
    a, b, c = (ol.fill("uniform_f64", m, seed=ol.SEED + k, first=lo) for k in range(3))  # counter-based: shard == slice
    da, db, dc = (ab.alloc_buf(dev, np.float64, m, q) for _ in range(3))
    for d, h in ((da, a), (db, b), (dc, c)):
        ab.memcpy(q, d, h)
    ab.babelstream.triad(q, da, db, dc, 2.0)
    got = np.empty(m)
    ab.memcpy(q, got, dc)
    q.wait()
    fa, fb, fc = (ol.fill("uniform_f64", n, seed=ol.SEED + k) for k in range(3))
    assert a.tobytes() == fa[lo:hi].tobytes(), "counter-based fill is not shard-consistent"
    ol.orc_stream("triad", fa, fb, fc, scalar=2.0)
    assert got.tobytes() == fc[lo:hi].tobytes(), f"rank {rank}: sharded Triad differs from the unsharded oracle"

    d_local = float(ab.babelstream.dot(q, da, db))
    d_total = multi.dot_all_ranks(d_local, dist, torch.device("cuda", local_rank))
    fa, fb = (ol.fill("uniform_f64", n, seed=ol.SEED + k) for k in range(2))
    d_orc = ol.oracle().orc_dot_f64(P(fa), P(fb), n, 256, 1024, None)
    assert abs(d_total - d_orc) <= 1e-12 * float(np.sum(np.abs(fa * fb))), "sharded Dot outside 1e-12"

    # ---- slab-sharded uint32 reduce (wrap-add is order-free)
    x = ol.fill("hash_u32", m, seed=9, first=lo)
    dx_ = ab.alloc_buf(dev, np.uint32, m, q)
    ab.memcpy(q, dx_, x)
    r_local = int(ab.reduce.reduce_sum(q, dx_))
    t = torch.tensor([r_local], dtype=torch.int64, device=f"cuda:{local_rank}")
    dist.all_reduce(t)
    full = ol.fill("hash_u32", n, seed=9)
    assert int(t.item()) % 2**32 == int(full.astype(np.uint64).sum()) % 2**32, "sharded reduce differs"

    # ---- the same Dot and reduce with the exchange FUSED into the reduction launch (peer stores + flags, no NCCL)
    ex = multi.ScalarExchange(q, rank, world)
    multi.connect_exchange_over_process_group(ex, dist)
    for _ in range(3):  # repeated collective calls (slot double-buffering)
        d_fused = float(ex.dot(q, da, db))
        assert d_fused == d_total, f"fused Dot exchange {d_fused} differs from the rank-ordered combination {d_total}"
        r_fused = int(ex.reduce_sum(q, dx_))
        assert r_fused == int(t.item()) % 2**32, "fused reduce exchange differs"
    assert ex.status() == 0
    dist.barrier()
    ex.close()

    # ---- decomposed heat, IPC peer pointers, fused halo exchange
    py, px = decomp.process_grid(world)
    NY, NX, steps = 192 * py, 640 * px, 25
    tile = decomp.tile_for(rank, world, NY, NX)
    runner = multi.HeatTile(q, tile, NY, NX)
    multi.connect_over_process_group(runner, dist)
    runner.upload(runner.initial_field())
    dist.barrier()
    runner.step(steps)
    q.wait()
    assert runner.status() == 0, f"rank {rank}: halo flag wait timed out"
    local = runner.download()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((tile, local), gathered, dst=0)
    if rank == 0:
        out = np.full((NY + 2, NX + 2), np.nan)
        for tl, f in gathered:
            decomp.stitch(out, tl, f)
        dx, dy, dt = ol.heat_params(NY, NX)
        u0 = np.empty((NY + 2, NX + 2))
        ol.oracle().orc_heat2d_init(P(u0), NY, NX, NX + 2, dx, dy)
        want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
        mask = np.ones_like(want, dtype=bool)
        mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = False
        assert out[mask].tobytes() == want[mask].tobytes(), "decomposed heat differs from the undecomposed oracle"
    dist.barrier()
    runner.close()

    # ---- row slabs, 2, 3, 4, 6 and 8 time levels per launch and per exchange (ghost rows that deep), IPC peer pointers
    NYs, NXs, steps2 = 160 * world, 900, 24
    dxs, dys, dts = ol.heat_params(NYs, NXs)
    u0s = np.empty((NYs + 2, NXs + 2))
    ol.oracle().orc_heat2d_init(P(u0s), NYs, NXs, NXs + 2, dxs, dys)
    want_s = ol.orc_heat_run(u0s, 1, steps2, dxs, dys, dts) if rank == 0 else None
    for levels in (2, 3, 4, 6, 8):
        slab = multi.HeatSlab(q, rank, world, NYs, NXs, levels=levels)
        multi.connect_over_process_group(slab, dist)
        slab.upload(slab.window(u0s))
        dist.barrier()
        slab.step(steps2)
        q.wait()
        assert slab.status() == 0, f"rank {rank}: slab flag wait timed out"
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((slab.g0, slab.owned_rows(), slab.download()), gathered, dst=0)
        if rank == 0:
            out = np.full((NYs + 2, NXs + 2), np.nan)
            for g0, (j0, j1), f in gathered:
                out[g0 + j0 : g0 + j1, :] = f[j0:j1, :]
            assert out.tobytes() == want_s.tobytes(), f"slab-decomposed {levels}-level heat differs from the undecomposed oracle"
        dist.barrier()
        slab.close()
    # ---- 2-D tiles, 4 and 8 time levels per launch (ghost cells that deep on all sides; rows inside the walker launch, columns
    # and corners by the column kernel), IPC peer pointers
    NYd, NXd, steps3 = 96 * py, 320 * px, 24
    dxd, dyd, dtd = ol.heat_params(NYd, NXd)
    u0d = ol.fill("uniform_f64", (NYd + 2) * (NXd + 2), seed=77).reshape(NYd + 2, NXd + 2)
    want_d = ol.orc_heat_run(u0d, 1, steps3, dxd, dyd, dtd) if rank == 0 else None
    for levels in (4, 8):
        deep = multi.HeatTileDeep(q, rank, world, NYd, NXd, levels=levels)
        multi.connect_over_process_group(deep, dist)
        deep.upload(deep.window(u0d))
        dist.barrier()
        deep.step(steps3)
        q.wait()
        assert deep.status() == 0, f"rank {rank}: deep tile flag wait timed out"
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((deep.tile, deep.download()), gathered, dst=0)
        if rank == 0:
            out = np.full((NYd + 2, NXd + 2), np.nan)
            for tl, f in gathered:
                tl.stitch(out, f)
            mask = np.ones_like(want_d, dtype=bool)
            mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = False
            assert out[mask].tobytes() == want_d[mask].tobytes(), f"2-D tiles at {levels} levels per launch differ from the undecomposed oracle"
        dist.barrier()
        deep.close()
    # ---- every sharded path against the committed golden vectors of the UNMODIFIED reference (tests/golden_multi.py;
    # the same block runs inside bench.py's N > 1 arm)
    import golden_multi

    verdict = golden_multi.check_all(golden_multi.Ranks(ab, world, {rank: q}, dist))
    if rank == 0:
        print(f"MP_WORKER_GOLDEN {verdict}", flush=True)
        print(f"MP_WORKER_OK {world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
