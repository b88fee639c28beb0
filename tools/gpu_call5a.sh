#!/bin/bash
# round-2 GPU call 5a (1 GPU): Mul/Add launch-shape sweep (Mul sits 3.5 % under Copy at the same shape), and the
# small-field regime of the 4-level heat launch (2048 x 16384: one GPU's share at 8 GPUs) without any exchange
O=gpurun_out/r02; mkdir -p $O
python - <<'PY' > $O/tune_mul.log 2>&1
import sys, itertools, numpy as np
sys.path.insert(0, '.')
import alpaka_b200 as ab
dev = ab.Platform().get_dev_by_idx(0); q = ab.Queue(dev)
n = 1 << 30
bs = ab.babelstream
a, b, c = (ab.alloc_buf(dev, np.float64, n, q) for _ in range(3))
bs.init(q, a, b, c)
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
def t(fn, k=20):
    for _ in range(3): fn()
    q.wait(); ab.enqueue(q, e0)
    for _ in range(k): fn()
    ab.enqueue(q, e1); q.wait(); return e0.elapsed_ms(e1) / k
runs = {"copy a->c": (lambda: bs.copy(q, a, c), 16.0, "copy"), "mul c->b": (lambda: bs.mul(q, c, b), 16.0, "mul"),
        "mul a->c": (lambda: bs.mul(q, a, c), 16.0, "mul"), "copy c->b": (lambda: bs.copy(q, c, b), 16.0, "copy"),
        "add": (lambda: bs.add(q, a, b, c), 24.0, "add"), "triad": (lambda: bs.triad(q, a, b, c), 24.0, "triad")}
for name, (fn, bpe, op) in runs.items():
    shapes = [(4, 256), (1, 1024), (2, 512), (4, 128), (2, 256), (1, 512)]
    for (u, blk), hint in itertools.product(shapes, (1, 2, 3, 0)):
        for k, v in (("unroll", u), ("block", blk), ("hint", hint)):
            ab.runtime.tune_set(f"stream.{op}.{k}", v)
        ms = t(fn)
        print(f"{name:10s} unroll={u} block={blk:4d} hint={hint}: {ms:.4f} ms {bpe * n * 1e-6 / ms:8.1f} GB/s", flush=True)
PY
echo "tune_mul rc=$?"; sort -k8 -n -r $O/tune_mul.log | head -5
python - <<'PY' > $O/placement.log 2>&1
# does the distance between the array read and the array written move Mul/Copy? (a->b sits 3.5 % under a->c)
import sys, numpy as np
sys.path.insert(0, '.')
import alpaka_b200 as ab
dev = ab.Platform().get_dev_by_idx(0); q = ab.Queue(dev)
n = 1 << 30
bs = ab.babelstream
pad = 1 << 24  # doubles (128 MiB)
big = ab.alloc_buf(dev, np.float64, 3 * n + 4 * pad, q)
check = ab.alloc_buf(dev, np.float64, 16, q)
def view(off): return ab.runtime.Buf(dev, np.float64, n, native_ptr=big.ptr + 8 * off)
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
def t(fn, k=20):
    for _ in range(3): fn()
    q.wait(); ab.enqueue(q, e0)
    for _ in range(k): fn()
    ab.enqueue(q, e1); q.wait(); return e0.elapsed_ms(e1) / k
src = view(0)
bs.init(q, src, view(n), view(2 * n))
for dist_elems in (n, n + 512, n + 4096, n + (1 << 15), n + (1 << 17), n + (1 << 18), n + (1 << 19), n + (1 << 20), n + (1 << 21), n + (1 << 22), n + (1 << 23), n + pad, 2 * n, 2 * n + (1 << 20), n + n // 2, n + n // 4, n + 3 * (1 << 20)):
    dst = view(dist_elems)
    for op, fn in (("copy", lambda: bs.copy(q, src, dst)), ("mul", lambda: bs.mul(q, src, dst))):
        ms = t(fn)
        print(f"{op:5s} dst - src = {dist_elems * 8 / 2**30:9.5f} GiB: {ms:.4f} ms {16.0 * n * 1e-6 / ms:8.1f} GB/s", flush=True)
PY
echo "placement rc=$?"; cat $O/placement.log
python - <<'PY' > $O/heat_small_field.log 2>&1
import sys, numpy as np
sys.path.insert(0, '.')
import alpaka_b200 as ab
dev = ab.Platform().get_dev_by_idx(0); q = ab.Queue(dev)
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
for NY, NX in ((16384, 16384), (8192, 16384), (4096, 16384), (2048, 16384)):
    dx = dy = 1.0 / (NX + 1); dt = 0.2 * dx * dx
    h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
    h.upload(ab.heat2d.initial_field(NY, NX, dx, dy))
    for launches in (25, 250):
        h.step(20, fuse=4); q.wait(); ab.enqueue(q, e0)
        h.step(4 * launches, fuse=4)
        ab.enqueue(q, e1); q.wait()
        us = e0.elapsed_ms(e1) / launches * 1e3
        print(f"heat {NY}x{NX} 4 levels/launch, {launches} launches: {us:.1f} us/launch = {us * 16384 / NY:.1f} us scaled to 16384 rows", flush=True)
    h.close()
PY
echo "heat_small rc=$?"; cat $O/heat_small_field.log
