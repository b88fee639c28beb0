"""tools/heat_step2_probe.py [ny nx] -- the two-levels-per-launch heat kernel (b200_heat2d_step2_f64) against the one-step
kernel on ONE GPU: every tile shape (heat.step2_ty x heat.step2_rpt), time per launch and per step, algorithmic GB/s
(16 B per cell per step). A small rough field is checked bit for bit between the two paths first."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import alpaka_b200 as ab


def timed(q, dev, fn, steps=100, warm=10):
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
    for _ in range(warm):
        fn()
    q.wait()
    ab.enqueue(q, e0)
    for _ in range(steps):
        fn()
    ab.enqueue(q, e1)
    q.wait()
    return e0.elapsed_ms(e1) / steps


def main():
    ny, nx = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
    dev = ab.Platform().get_dev_by_idx(0)
    q = ab.Queue(dev)

    # parity between the two device paths on a rough field (the oracle comparison lives in tests/test_gpu_heat2d.py)
    sy_, sx_ = 203, 391
    dx, dy = 1.0 / (sx_ + 1), 1.0 / (sy_ + 1)
    dt = 0.2 * min(dx * dx, dy * dy)
    rng = np.random.default_rng(5)
    u0 = rng.uniform(-1, 1, (sy_ + 2, sx_ + 2))
    outs = []
    for fuse in (1, 2, 3, 4):
        h = ab.heat2d.Heat2D(q, sy_, sx_, dx, dy, dt)
        h.upload(u0)
        h.step(9, fuse=fuse)
        outs.append(h.download())
        h.close()
    print("1 / 2 / 3 / 4 levels per launch, 9 steps, 203x391 rough field: bit-identical =",
          all(o.tobytes() == outs[0].tobytes() for o in outs[1:]))

    dx, dy = 1.0 / (nx + 1), 1.0 / (ny + 1)
    dt = 0.2 * min(dx * dx, dy * dy)
    h = ab.heat2d.Heat2D(q, ny, nx, dx, dy, dt)
    h.upload(ab.heat2d.initial_field(ny, nx, dx, dy))  # real data: power draw depends on the operand bits
    ms1 = timed(q, dev, lambda: h.step(1))
    print(f"one step per launch            {ny}x{nx}: {ms1 * 1e3:8.1f} us/step  {16.0 * ny * nx * 1e-9 / (ms1 * 1e-3):8.1f} GB/s")
    for ty, rpt in ((32, 8), (32, 16), (32, 32), (64, 16), (64, 32)):
        ab.runtime.tune_set("heat.step2_ty", ty)
        ab.runtime.tune_set("heat.step2_rpt", rpt)
        ms2 = timed(q, dev, lambda: h.step(2))
        print(f"two steps per launch ty={ty:2d} rpt={rpt:2d} {ny}x{nx}: {ms2 * 1e3 / 2:8.1f} us/step  "
              f"{2 * 16.0 * ny * nx * 1e-9 / (ms2 * 1e-3):8.1f} GB/s algorithmic  ({ms2 * 1e3:.1f} us/launch, x{2 * ms1 / ms2:.2f})")
    for levels, rpt, nwy in ((3, 16, 4), (3, 16, 2), (3, 32, 2), (4, 16, 4), (4, 32, 2)):
        ab.runtime.tune_set("heat.stepn_rpt", rpt)
        ab.runtime.tune_set("heat.stepn_nwy", nwy)
        msn = timed(q, dev, lambda: h.step(levels, fuse=levels))
        print(f"{levels} steps per launch rpt={rpt:2d} nwy={nwy} {ny}x{nx}: {msn * 1e3 / levels:8.1f} us/step  "
              f"{levels * 16.0 * ny * nx * 1e-9 / (msn * 1e-3):8.1f} GB/s algorithmic  ({msn * 1e3:.1f} us/launch, x{levels * ms1 / msn:.2f})")
    # sustained: 1000 steps in one go (BASELINE.json's heat configs); the power cap decides here, not the burst figure
    print("sustained, 1000 steps per measurement (after 1000 warm-up steps):")
    for levels, rpt, nwy in ((1, 0, 0), (2, 0, 0), (3, 16, 2), (3, 16, 4), (3, 32, 2), (4, 16, 4)):
        if levels >= 3:
            ab.runtime.tune_set("heat.stepn_rpt", rpt)
            ab.runtime.tune_set("heat.stepn_nwy", nwy)
        if levels == 2:
            ab.runtime.tune_set("heat.step2_ty", 64)
            ab.runtime.tune_set("heat.step2_rpt", 16)
        n = 1000 - 1000 % levels
        ms = timed(q, dev, lambda: h.step(n, fuse=levels), steps=1, warm=1) / n
        print(f"  {levels} level(s) per launch rpt={rpt:2d} nwy={nwy}: {ms * 1e3:8.1f} us/step  {16.0 * ny * nx * 1e-9 / (ms * 1e-3):8.1f} GB/s algorithmic")
    h.close()


main()
