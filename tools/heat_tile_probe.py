"""tools/heat_tile_probe.py -- per-step time of one heat tile: plain fused step vs the five-window (strip + interior)
launch used by the decomposed run, on ONE GPU (no neighbours). Separates launch-shape overhead from exchange cost."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import alpaka_b200 as ab
from alpaka_b200 import decomp, multi

def timed(q, dev, fn, steps=200, warm=20):
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
    for _ in range(warm): fn()
    q.wait(); ab.enqueue(q, e0)
    for _ in range(steps): fn()
    ab.enqueue(q, e1); q.wait()
    return e0.elapsed_ms(e1) / steps

def main():
    ny, nx = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 8192)
    dev = ab.Platform().get_dev_by_idx(0); q = ab.Queue(dev)
    dx, dy = 1.0/(nx+1), 1.0/(ny+1); dt = 0.2*min(dx*dx, dy*dy)
    h = ab.heat2d.Heat2D(q, ny, nx, dx, dy, dt)
    h.upload(np.zeros((ny+2, nx+2)))
    ms = timed(q, dev, lambda: h.step(1))
    ideal = 16.0*ny*nx/6544.3e9*1e3
    print(f"plain fused step      {ny}x{nx}: {ms*1e3:8.2f} us/step  ({16.0*ny*nx*1e-9/(ms*1e-3):7.1f} GB/s, ideal {ideal*1e3:.1f} us)")
    h.close()
    tile = decomp.tile_for(0, 1, ny, nx)
    r = multi.HeatTile(q, tile, ny, nx); multi.connect_in_process([r]); r.upload(np.zeros((ny+2, nx+2)))
    ms = timed(q, dev, lambda: r.step(1))
    print(f"five-window, 0 peers  {ny}x{nx}: {ms*1e3:8.2f} us/step  ({16.0*ny*nx*1e-9/(ms*1e-3):7.1f} GB/s)")
    r.close()

main()
