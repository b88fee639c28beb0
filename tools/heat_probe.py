#!/usr/bin/env python
"""Run a few heat2d steps at 16384^2 (for ncu captures with B200_TUNE overrides)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alpaka_b200 as ab
from alpaka_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = ab.Platform().get_dev_by_idx(0)
q = ab.Queue(dev)
dx = dy = 1.0 / (n + 1)
h = ab.heat2d.Heat2D(q, n, n, dx, dy, 0.2 * dx * dx)
for b in h.bufs:
    _lib.load().b200_memset2d_async(dev.idx, b.ptr, b.pitch_bytes, 0, (n + 2) * 8, n + 2, q.handle)
h.step(steps)
q.wait()
print("ok")
