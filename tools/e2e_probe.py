"""tools/e2e_probe.py -- end-to-end Triad on pinned HOST arrays: the chunked copy pipeline (bench.py's e2e) against the
same kernel reading and writing the host arrays directly over PCIe (zero copy), and pipeline parameters."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import alpaka_b200 as ab
from alpaka_b200 import _lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
    dev = ab.Platform().get_dev_by_idx(0)
    q = ab.Queue(dev)
    lib = _lib.load()
    ha, hb, hc = (ab.alloc_mapped_buf(np.float64, n) for _ in range(3))
    ha.array[:] = 1.0
    hb.array[:] = 2.0

    def report(name, fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        t = (time.perf_counter() - t0) / reps
        assert float(hc.array[0]) == 5.0 and float(hc.array[-1]) == 5.0
        print(f"{name:58s} {24.0 * n * 1e-9 / t:7.1f} GB/s  ({t * 1e3:.1f} ms, H2D {16.0 * n * 1e-9 / t:.1f} GB/s)")

    for chunk, depth in ((1 << 23, 4), (1 << 22, 4), (1 << 24, 4), (1 << 23, 2), (1 << 23, 8), (1 << 21, 8)):
        pipe = ab.babelstream.TriadHostPipeline(dev, np.float64, chunk_elems=chunk, depth=depth)
        report(f"copy pipeline chunk=2^{chunk.bit_length() - 1} depth={depth}", lambda: pipe.run(ha.array, hb.array, hc.array))
        pipe.close()

    def zero_copy():
        hc.array[0] = 0.0
        hc.array[-1] = 0.0
        rc = lib.b200_stream_triad_f64(q.handle, ha.ptr, hb.ptr, hc.ptr, 2.0, n)
        assert rc == 0
        q.wait()

    report("zero copy: the Triad kernel on the pinned host pointers", zero_copy)


main()
