"""tools/heat_multi_probe.py -- torchrun worker: per-rank step time of the decomposed heat run for one tile size, with
the halo_debug knobs (1 = no peer stores, 2 = no flag wait) to attribute the exchange cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import alpaka_b200 as ab
from alpaka_b200 import decomp, multi

def main():
    ny, nx = int(sys.argv[1]), int(sys.argv[2])
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = ab.Platform().get_dev_by_idx(lr); q = ab.Queue(dev)
    py, px = decomp.process_grid(world)
    NY, NX = ny * py, nx * px
    tile = decomp.tile_for(rank, world, NY, NX)
    r = multi.HeatTile(q, tile, NY, NX); multi.connect_over_process_group(r, dist); r.upload(r.initial_field())
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
    for dbg in (0, 1, 2, 3, 0):
        ab.runtime.tune_set("heat.halo_debug", dbg)
        q.wait(); dist.barrier()
        for _ in range(5): r.step(1)
        q.wait(); dist.barrier(); ab.enqueue(q, e0)
        for _ in range(40): r.step(1)
        ab.enqueue(q, e1); q.wait()
        ms = e0.elapsed_ms(e1) / 40
        all_ms = [None] * world; dist.all_gather_object(all_ms, ms)
        if rank == 0:
            print(f"tile {ny}x{nx} grid {py}x{px} halo_debug={dbg}: per-rank us/step " + " ".join(f"{m*1e3:.1f}" for m in all_ms), flush=True)
        # re-sync the flag protocol after a debug phase (levels stay monotonic; ghosts are garbage in debug phases)
    r.close(); dist.destroy_process_group()

main()
