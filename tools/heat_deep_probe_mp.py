"""tools/heat_deep_probe_mp.py -- torchrun worker: deep 2-D tiles against row slabs on the same field, per-launch time with the
exchange switched off piecewise (heat.halo_debug: 1 = no peer stores, 2 = no flag waits, 4 = no column kernel).

    torchrun --nproc-per-node N tools/heat_deep_probe_mp.py NY NX PY PX [levels]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import alpaka_b200 as ab
from alpaka_b200 import multi


def main():
    NY, NX, PY, PX = (int(x) for x in sys.argv[1:5])
    G = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = ab.Platform().get_dev_by_idx(lr)
    q = ab.Queue(dev)
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)

    def timed(obj, n):
        q.wait(); dist.barrier()
        for _ in range(5): obj.step(G)
        q.wait(); dist.barrier(); ab.enqueue(q, e0)
        for _ in range(n): obj.step(G)
        ab.enqueue(q, e1); q.wait()
        out = [None] * world
        dist.all_gather_object(out, e0.elapsed_ms(e1) / n)
        return out

    for name, make in (("slabs", lambda: multi.HeatSlab(q, rank, world, NY, NX, levels=G)),
                       (f"deep tiles {PY}x{PX}", lambda: multi.HeatTileDeep(q, rank, world, NY, NX, levels=G, grid=(PY, PX)))):
        obj = make()
        multi.connect_over_process_group(obj, dist)
        obj.upload(obj.initial_field())
        for dbg in (0, 1, 2, 3, 6, 7, 0):
            ab.runtime.tune_set("heat.halo_debug", dbg)
            ms = timed(obj, 40)
            if rank == 0:
                print(f"{name:18s} {NY}x{NX} levels={G} halo_debug={dbg}: per-rank us/launch " + " ".join(f"{m * 1e3:.1f}" for m in ms), flush=True)
        ab.runtime.tune_set("heat.halo_debug", 0)
        obj.close()
    dist.destroy_process_group()


main()
