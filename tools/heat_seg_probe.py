"""tools/heat_seg_probe.py -- rows per walker segment against time per step on small fields (one GPU's share of a 16384^2 field
at 8 and 4 GPUs), four levels per launch. python tools/heat_seg_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alpaka_b200 as ab

dev = ab.Platform().get_dev_by_idx(0)
q = ab.Queue(dev)
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
NX = 16384
for NY in (2048, 4096):
    dx = dy = 1.0 / (NX + 1)
    dt = 0.2 * dx * dx
    h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
    h.upload(ab.heat2d.initial_field(NY, NX, dx, dy))
    for seg in (0, 40, 48, 54, 60, 64, 72, 79, 82, 90, 96, 103, 110, 128, 137, 160, 171, 205, 256):
        ab.runtime.tune_set("heat.walk_seg_rows", seg)
        best = 1e9
        for rep in range(3):
            h.step(16, fuse=4); q.wait(); ab.enqueue(q, e0)
            h.step(960, fuse=4)
            ab.enqueue(q, e1); q.wait()
            best = min(best, e0.elapsed_ms(e1) / 960 * 1e3)
        print(f"heat {NY}x{NX} seg_rows={seg:3d}: {best:6.2f} us/step", flush=True)
    ab.runtime.tune_set("heat.walk_seg_rows", 0)
    h.close()
