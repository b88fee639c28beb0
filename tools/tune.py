#!/usr/bin/env python
"""Parameter sweeps on the GPU box (development tool; prints one line per configuration).

    python tools/tune.py heat   [--n 16384]
    python tools/tune.py stream [--n 1073741824] [--kernels triad,copy]
    python tools/tune.py reduce
"""
import argparse
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import alpaka_b200 as ab  # noqa: E402
from alpaka_b200 import _lib  # noqa: E402


def timed(q, dev, fn, steps=10, warmup=3):
    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
    for _ in range(warmup):
        fn()
    q.wait()
    ab.enqueue(q, e0)
    for _ in range(steps):
        fn()
    ab.enqueue(q, e1)
    q.wait()
    return e0.elapsed_ms(e1) / steps


def sweep_heat(args, dev, q):
    ny = nx = args.n or 16384
    dx, dy = 1.0 / (nx + 1), 1.0 / (ny + 1)
    dt = 0.2 * min(dx * dx, dy * dy)
    for promo in (256, 128, 0):
        ab.runtime.tune_set("heat.l2promo", promo)
        h = ab.heat2d.Heat2D(q, ny, nx, dx, dy, dt)
        lib = _lib.load()
        for b in h.bufs:
            lib.b200_memset2d_async(dev.idx, b.ptr, b.pitch_bytes, 0, (nx + 2) * 8, ny + 2, q.handle)
        for rpt, stages, ctas, hint in itertools.product((4, 8), (2, 3, 4), (1, 2, 3), (0, 1)):
            if stages * 35968 * ctas > 227 * 1024:
                continue
            for k, v in (("heat.rpt", rpt), ("heat.stages", stages), ("heat.ctas_per_sm", ctas), ("heat.hint", hint)):
                ab.runtime.tune_set(k, v)
            try:
                ms = timed(q, dev, lambda: h.step(1), 20, 3)
            except ab.B200Error as e:
                print("heat", promo, rpt, stages, ctas, hint, "ERR", e)
                continue
            print(f"heat promo={promo} rpt={rpt} stages={stages} ctas={ctas} hint={hint}: {ms:.4f} ms  "
                  f"{16.0 * ny * nx * 1e-6 / ms:.1f} GB/s", flush=True)
        h.close()


def sweep_stream(args, dev, q):
    n = args.n or (1 << 30)
    bs = ab.babelstream
    a, b, c = (ab.alloc_buf(dev, np.float64, n, q) for _ in range(3))
    bs.init(q, a, b, c)
    runs = {
        "copy": (lambda: bs.copy(q, a, c), 16.0),
        "mul": (lambda: bs.mul(q, a, b), 16.0),
        "add": (lambda: bs.add(q, a, b, c), 24.0),
        "triad": (lambda: bs.triad(q, a, b, c), 24.0),
        "nstream": (lambda: bs.nstream(q, c, a, b, 0.0), 32.0),
        "init": (lambda: bs.init(q, a, b, c), 24.0),
    }
    names = args.kernels.split(",") if args.kernels else ["triad", "copy"]
    for name in names:
        fn, bpe = runs[name]
        for vb, unroll, hint, block, ctas in itertools.product((32,), (1, 2, 4), (0, 1, 2, 3), (256, 512), (0, 2, 4, 8)):
            if block * ctas > 2048:
                continue
            for k, v in (("vb", vb), ("unroll", unroll), ("hint", hint), ("block", block), ("ctas_per_sm", ctas)):
                ab.runtime.tune_set(f"stream.{k}", v)
            ms = timed(q, dev, fn, 8, 2)
            print(f"{name} vb={vb} unroll={unroll} hint={hint} block={block} ctas={ctas}: {ms:.4f} ms "
                  f"{bpe * n * 1e-6 / ms:.1f} GB/s", flush=True)


def sweep_reduce(args, dev, q):
    n = args.n or (1 << 30)
    bs = ab.babelstream
    a, b, c = (ab.alloc_buf(dev, np.float64, n, q) for _ in range(3))
    bs.init(q, a, b, c)
    out = ab.alloc_buf(dev, np.float64, 1, q)
    for unroll, ctas in itertools.product((1, 2, 4), (1, 2, 3, 4)):
        ab.runtime.tune_set("dot.unroll", unroll)
        ab.runtime.tune_set("dot.ctas_per_sm", ctas)
        ms = timed(q, dev, lambda: bs.dot_async(q, a, b, out), 10, 3)
        print(f"dot unroll={unroll} ctas={ctas}: {ms:.4f} ms {16.0 * n * 1e-6 / ms:.1f} GB/s", flush=True)
    src = ab.create_view(dev, a.ptr, np.uint32, 2 * n)
    res = ab.alloc_buf(dev, np.uint32, 1, q)
    for unroll, ctas in itertools.product((1, 2, 4), (1, 2, 3, 4)):
        ab.runtime.tune_set("reduce.unroll", unroll)
        ab.runtime.tune_set("reduce.ctas_per_sm", ctas)
        ms = timed(q, dev, lambda: ab.reduce.reduce_sum_async(q, src, res), 10, 3)
        print(f"reduce_u32 unroll={unroll} ctas={ctas}: {ms:.4f} ms {8.0 * n * 1e-6 / ms:.1f} GB/s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["heat", "stream", "reduce"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--kernels", default="")
    args = ap.parse_args()
    dev = ab.Platform().get_dev_by_idx(0)
    q = ab.Queue(dev)
    {"heat": sweep_heat, "stream": sweep_stream, "reduce": sweep_reduce}[args.what](args, dev, q)


if __name__ == "__main__":
    main()
