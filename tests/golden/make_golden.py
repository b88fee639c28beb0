#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates tests/golden/reference_vectors.npz from the UNMODIFIED reference.

Runs in the build container only (needs /root/reference to build oracle/_ref/libalpaka_ref.so, see oracle/Makefile):
the reference's own kernel functors are executed on its own CPU back-ends (AccCpuOmp2Blocks unless noted) on small
seeded inputs and the inputs + outputs are stored. The fixture travels to the GPU box, where /root/reference does
not exist; tests/test_golden.py checks the C oracle (CPU) and the CUDA path (GPU) against it bit for bit.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from oracle_lib import P  # noqa: E402


def main():
    ref = ol.ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libalpaka_ref.so is not available: build it with `make -C oracle ref`")
    out = {}
    rng = np.random.default_rng(20261017)

    # ---- BabelStream: every kernel once from the same initial state, both precisions
    n = 1000
    for dtype, tag in ((np.float64, "f64"), (np.float32, "f32")):
        a0 = rng.uniform(-1, 1, n).astype(dtype)
        b0 = rng.uniform(-1, 1, n).astype(dtype)
        c0 = rng.uniform(-1, 1, n).astype(dtype)
        out[f"stream_{tag}_in"] = np.stack([a0, b0, c0])
        for k in ("init", "copy", "mul", "add", "triad", "nstream"):
            a, b, c = a0.copy(), b0.copy(), c0.copy()
            ol.ref_stream(k, a, b, c, acc=1, init_a=1.0)
            out[f"stream_{tag}_{k}"] = np.stack([a, b, c])
        # Dot: reference DotKernel on the CPU back-end, WorkDiv {G,1,1}, host std::reduce of the partials
        for grid in (1, 7, 256):
            partials = np.empty(grid, dtype=dtype)
            d = ref.ref_babelstream_dot(1, 1 if dtype == np.float64 else 0, P(a0), P(b0), n, grid, P(partials))
            out[f"dot_{tag}_g{grid}"] = np.array([d], dtype=dtype)
            out[f"dot_{tag}_g{grid}_partials"] = partials

    # ---- example/reduce: reference ReduceKernel launched twice, AccCpuSerial (one block: fixed order)
    for dtype, tag in ((np.uint32, "u32"), (np.int32, "i32"), (np.uint64, "u64"), (np.float32, "f32"), (np.float64, "f64")):
        for m in (1, 17, 1000, 20011):
            if np.dtype(dtype).kind == "f":
                x = rng.integers(0, 2, m).astype(dtype)  # {0,1}: every partial sum is exact in any order
            else:
                x = rng.integers(0, 2**31 - 1, m).astype(dtype)
            out[f"reduce_{tag}_n{m}_in"] = x
            out[f"reduce_{tag}_n{m}"] = np.array([ol.ref_reduce(x, 0)], dtype=dtype)

    # ---- heatEquation2D: reference Stencil + Boundary kernels, driver loop
    for ny, nx, steps in ((16, 16, 100), (32, 48, 25)):
        dx, dy, dt = ol.heat_params(ny, nx)
        u = np.empty((ny + 2, nx + 2))
        ref.ref_heat2d_init(P(u), ny, nx, dx, dy)
        out[f"heat_{ny}x{nx}_init"] = u.copy()
        assert ref.ref_heat2d_run(1, P(u), ny, nx, 1, steps, dx, dy, dt, None) == 0
        out[f"heat_{ny}x{nx}_s{steps}"] = u
        out[f"heat_{ny}x{nx}_params"] = np.array([dx, dy, dt, steps])

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
