"""tools/heat_one.py -- a few fused heat launches on one GPU, for ncu: python tools/heat_one.py LEVELS [NY NX [launches]]
(tunables through B200_TUNE=...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alpaka_b200 as ab

S = int(sys.argv[1])
NY, NX = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16384, 16384)
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = ab.Platform().get_dev_by_idx(0)
q = ab.Queue(dev)
dx = dy = 1.0 / (NX + 1)
dt = 0.2 * dx * dx
h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
h.upload(ab.heat2d.initial_field(NY, NX, dx, dy))
h.step(S * launches, fuse=S)
q.wait()
h.close()
print("heat_one ok")
