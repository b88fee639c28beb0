import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import alpaka_b200 as ab
dev = ab.Platform().get_dev_by_idx(0); q = ab.Queue(dev)
ny = nx = 16384
dx, dy = 1.0/(nx+1), 1.0/(ny+1); dt = 0.2*min(dx*dx, dy*dy)
h = ab.heat2d.Heat2D(q, ny, nx, dx, dy, dt); h.upload(ab.heat2d.initial_field(ny, nx, dx, dy))
def timed(fn, steps, warm):
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
    for _ in range(warm): fn()
    q.wait(); ab.enqueue(q, e0)
    for _ in range(steps): fn()
    ab.enqueue(q, e1); q.wait()
    return e0.elapsed_ms(e1)/steps
for levels, rpt, nwy in ((3,16,2),(4,16,4),(4,16,2),(3,16,2),(4,16,4)):
    ab.runtime.tune_set("heat.stepn_rpt", rpt); ab.runtime.tune_set("heat.stepn_nwy", nwy)
    time.sleep(12)
    b = timed(lambda: h.step(levels, fuse=levels), 20, 5)/levels
    time.sleep(12)
    n = 1000 - 1000 % levels
    s_ = timed(lambda: h.step(n, fuse=levels), 1, 0)/n
    print(f"levels={levels} rpt={rpt} nwy={nwy}: cold burst {b*1e3:7.1f} us/step   cold 1000 steps {s_*1e3:7.1f} us/step", flush=True)
