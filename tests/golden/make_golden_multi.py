#!/usr/bin/env python
"""tests/golden/make_golden_multi.py -- regenerates tests/golden/multi_gpu_vectors.npz from the UNMODIFIED reference.

The fixture behind bench.py's `parity_n` block (N > 1) and tests/test_gpu_golden_multi.py: UNSHARDED inputs and the
outputs the reference's own CPU back-end (oracle/_ref, AccCpuOmp2Blocks / AccCpuSerial) produces for them, at sizes
that split over 2, 4 and 8 ranks. The sharded GPU paths must reproduce them: slab-sharded Triad / Nstream windows bit for
bit, the fused Dot / reduce exchange, the decomposed heat field (2-D tiles and row slabs, every launch depth) stitched
back together. Runs in the build container only (needs /root/reference for oracle/_ref).

    python tests/golden/make_golden_multi.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from oracle_lib import P  # noqa: E402

N_STREAM = (1 << 15) + 40  # ragged: the last slab is shorter, slab starts are 4-element aligned
HEAT = (128, 256, 24)      # NY, NX, steps: tiles 2x1 / 2x2 / 4x2, slabs of 64 / 32 / 16 rows; 24 = 12x2 = 8x3 = 6x4 launches


def main():
    ref = ol.ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libalpaka_ref.so is not available: build it with `make -C oracle ref`")
    out = {}
    rng = np.random.default_rng(20261018)
    n = N_STREAM

    # ---- BabelStream: Triad and Nstream once from the same state (scalar = 2, babelStreamCommon.hpp:31)
    a0, b0, c0 = (rng.uniform(-1, 1, n) for _ in range(3))
    out["stream_in"] = np.stack([a0, b0, c0])
    for k in ("triad", "nstream"):
        a, b, c = a0.copy(), b0.copy(), c0.copy()
        ol.ref_stream(k, a, b, c, acc=1)
        out[f"stream_{k}"] = c if k == "triad" else a
    # Dot: reference DotKernel on AccCpuOmp2Blocks, WorkDiv {256,1,1}, host std::reduce (babelStreamMainTest.cpp:402-403)
    out["dot_uniform"] = np.array([ref.ref_babelstream_dot(1, 1, P(a0), P(b0), n, 256, None)])
    # small integers: every product and partial sum is exact, so ANY summation order must give the same bits
    ia, ib = (rng.integers(-8, 9, n).astype(np.float64) for _ in range(2))
    out["dot_int_in"] = np.stack([ia, ib])
    out["dot_int"] = np.array([ref.ref_babelstream_dot(1, 1, P(ia), P(ib), n, 256, None)])
    assert out["dot_int"][0] == float(np.dot(ia.astype(np.int64), ib.astype(np.int64)))

    # ---- example/reduce: reference ReduceKernel twice on AccCpuSerial
    xu = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    out["reduce_u32_in"] = xu
    out["reduce_u32"] = np.array([ol.ref_reduce(xu, 0)], dtype=np.uint32)
    xf = rng.integers(0, 2, n).astype(np.float32)  # {0,1}: exact in any order
    out["reduce_f32_in"] = xf
    out["reduce_f32"] = np.array([ol.ref_reduce(xf, 0)], dtype=np.float32)

    # ---- heatEquation2D: the undecomposed field after `steps` steps of the reference's Stencil + Boundary kernels
    ny, nx, steps = HEAT
    dx, dy, dt = ol.heat_params(ny, nx)
    u = np.empty((ny + 2, nx + 2))
    ref.ref_heat2d_init(P(u), ny, nx, dx, dy)
    out["heat_init"] = u.copy()
    assert ref.ref_heat2d_run(1, P(u), ny, nx, 1, steps, dx, dy, dt, None) == 0
    out["heat_final"] = u
    out["heat_params"] = np.array([dx, dy, dt, steps])

    path = os.path.join(HERE, "multi_gpu_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
