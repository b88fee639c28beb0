"""tools/heat_slab_probe.py -- torchrun worker: the small-slab regime of the fused multi-level heat launch (what each
GPU sees at 8 GPUs on 16384^2: 2048 rows) reproduced on fewer GPUs. Per-launch time of `levels` time levels on NY/world
rows x NX columns per rank, burst and over 1000 steps, with the halo_debug knobs (1 = no peer stores, 2 = no flag wait)
to attribute the exchange cost.

    torchrun --nproc-per-node 2 tools/heat_slab_probe.py NY NX [levels]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import alpaka_b200 as ab
from alpaka_b200 import multi


def main():
    NY, NX = int(sys.argv[1]), int(sys.argv[2])
    G = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = ab.Platform().get_dev_by_idx(lr)
    q = ab.Queue(dev)
    s = multi.HeatSlab(q, rank, world, NY, NX, levels=G)
    multi.connect_over_process_group(s, dist)
    s.upload(s.initial_field())
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)

    def timed(n_launches):
        q.wait(); dist.barrier()
        s.step(G * 5)
        q.wait(); dist.barrier(); ab.enqueue(q, e0)
        s.step(G * n_launches)
        ab.enqueue(q, e1); q.wait()
        ms = e0.elapsed_ms(e1) / n_launches
        out = [None] * world
        dist.all_gather_object(out, ms)
        return out

    for dbg in (0, 1, 2, 3, 0):
        ab.runtime.tune_set("heat.halo_debug", dbg)
        for n in (25, 250):
            ms = timed(n)
            if rank == 0:
                print(f"slab {NY // world}x{NX} x{world} levels={G} halo_debug={dbg} launches={n}: per-rank us/launch "
                      + " ".join(f"{m * 1e3:.1f}" for m in ms), flush=True)
    ab.runtime.tune_set("heat.halo_debug", 0)
    assert s.status() == 0
    s.close()
    dist.destroy_process_group()


main()
