"""tools/heat_walk_probe.py -- the walker kernel (heatWalkKernel) against the tile kernel on one GPU: microseconds per step
in a short burst and over 1000 steps (sustained clocks), per depth and stage shape.

    python tools/heat_walk_probe.py [NY NX]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alpaka_b200 as ab

NY, NX = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
dev = ab.Platform().get_dev_by_idx(0)
q = ab.Queue(dev)
dx = dy = 1.0 / (NX + 1)
dt = 0.2 * dx * dx
h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
h.upload(ab.heat2d.initial_field(NY, NX, dx, dy))
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)


def t(fn, n):
    fn(); q.wait(); ab.enqueue(q, e0)
    for _ in range(n):
        fn()
    ab.enqueue(q, e1); q.wait()
    return e0.elapsed_ms(e1) / n


long_steps = int(os.environ.get("LONG_STEPS", "960"))  # a multiple of 2, 3, 4, 6, 8
RESET = {"heat.walk": 1, "heat.walk_shape": 0, "heat.walk_seg_rows": 0, "heat.walk_minb": 0, "heat.walk_lds_swap": 0, "heat.walk_split": 0,
         "heat.walk_pdl": 1}
cases = [("tile  4", 4, {"heat.walk": 0})]
for rep in range(2):
    for S in (4, 6, 8):
        cases.append((f"walk  {S} default (one kernel)", S, {}))
        if S != 8:
            cases.append((f"walk  {S} split", S, {"heat.walk_split": 1}))
            cases.append((f"walk  {S} split, no pdl", S, {"heat.walk_split": 1, "heat.walk_pdl": 0}))
        cases.append((f"walk  {S} lds_swap", S, {"heat.walk_lds_swap": 1}))
for seg in (64, 128, 256, 512):
    for S in (4, 6):
        cases.append((f"walk  {S} seg{seg}", S, {"heat.walk_seg_rows": seg}))
for name, S, tune in cases:
    for k, v in tune.items():
        ab.runtime.tune_set(k, v)
    try:
        burst = t(lambda: h.step(S, fuse=S), 25) / S
        long_ = t(lambda: h.step(long_steps, fuse=S), 1) / long_steps
        print(f"heat {NY}x{NX} {name:22s}: burst {burst * 1e3:7.1f} us/step, {long_steps} steps {long_ * 1e3:7.1f} us/step", flush=True)
    except ab.B200Error as e:
        print(f"heat {NY}x{NX} {name}: ERROR {e}", flush=True)
    for k in tune:
        ab.runtime.tune_set(k, RESET[k])
h.close()
