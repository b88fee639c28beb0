#!/usr/bin/env python
"""Dot with the all-ranks exchange fused into the reduction launch, TWO DEVICES IN ONE PROCESS (plain peer pointers, no
IPC): the form `ncu` can profile (one process; `--devices 0` captures rank 0's kernel, whose last block stores its scalar
into device 1's slot array and reads device 1's scalar out of its own) -- NVLink counters of reduceKernel<..., Exchange>.
Under ncu the profiled device's launches are serialised and replayed, so device 1 is enqueued FIRST (its kernel, not
profiled, publishes its scalar and then waits for device 0's, which the first replay pass delivers); the flag-wait bound is
lowered to 5 s so that a mistake here costs seconds of GPU time, not the 60 s default per replay pass.

    python tools/exchange_two_devices.py [n]        # n doubles per device (default 2^28)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import alpaka_b200 as ab  # noqa: E402
from alpaka_b200 import multi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
    platform = ab.Platform()
    if platform.get_dev_count() < 2:
        raise SystemExit("needs two devices")
    platform.enable_peer_access()
    devs = [platform.get_dev_by_idx(i) for i in range(2)]
    queues = [ab.Queue(d) for d in devs]
    exs = [multi.ScalarExchange(q, r, 2) for r, q in enumerate(queues)]
    multi.connect_exchange_in_process(exs)
    bufs, outs = [], []
    for d, q in zip(devs, queues):
        a, b, c = (ab.alloc_buf(d, np.float64, n, q) for _ in range(3))
        ab.babelstream.init(q, a, b, c)
        ab.babelstream.copy(q, a, b)
        ab.babelstream.mul(q, a, b)
        bufs.append((a, b))
        outs.append(ab.alloc_buf(d, np.float64, 1, q))
    ab.runtime.tune_set("exchange.timeout_ms", 5000)
    for call in range(6):
        for r in (1, 0):  # enqueue on both devices before waiting on either; the profiled device last (see above)
            exs[r].dot_async(queues[r], bufs[r][0], bufs[r][1], outs[r])
        got = []
        for r, q in enumerate(queues):
            h = np.empty(1)
            ab.memcpy(q, h, outs[r])
            q.wait()
            got.append(float(h[0]))
        assert got == [2.0 * n * 2] * 2, got
    for e in exs:
        assert e.status() == 0
    print(f"exchange_two_devices ok: dot over 2 x {n} doubles = {got[0]} on both devices")


if __name__ == "__main__":
    main()
