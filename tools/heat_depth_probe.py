"""tools/heat_depth_probe.py -- which depth (time levels per launch) is fastest at which field height, one GPU, default shapes:
us per step over a long run (sustained clocks). python tools/heat_depth_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alpaka_b200 as ab

dev = ab.Platform().get_dev_by_idx(0)
q = ab.Queue(dev)
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
NX = 16384
for NY in (1024, 2048, 4096, 8192, 16384):
    dx = dy = 1.0 / (NX + 1)
    dt = 0.2 * dx * dx
    h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
    h.upload(ab.heat2d.initial_field(NY, NX, dx, dy))
    steps = 960 if NY >= 8192 else 1920
    out = []
    for name, S, tune in (("tile4", 4, {"heat.walk": 0}), ("walk4", 4, {}), ("walk6", 6, {}), ("walk8", 8, {}), ("tile3", 3, {})):
        for k, v in tune.items():
            ab.runtime.tune_set(k, v)
        h.step(S * 4, fuse=S); q.wait(); ab.enqueue(q, e0)
        h.step(steps, fuse=S)
        ab.enqueue(q, e1); q.wait()
        out.append(f"{name} {e0.elapsed_ms(e1) / steps * 1e3:7.2f}")
        for k in tune:
            ab.runtime.tune_set(k, 1)
    print(f"heat {NY:5d}x{NX} us/step over {steps} steps: " + "  ".join(out), flush=True)
    h.close()
