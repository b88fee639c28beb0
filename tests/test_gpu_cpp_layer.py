"""GPU tests of the C++ alpaka API layer (include/alpaka) through the driver binaries built by examples/Makefile.

What is proven here:
  * the reference's OWN babelstream and heatEquation2D driver translation units, compiled unmodified against
    include/alpaka, run and self-validate on the B200 accelerator (Dot included: TagGpuCudaRt names the B200 back-end);
  * the same kernels give bit-identical results to the CPU oracle both through the hand-written kernels (functor
    recognition, alpaka/b200/Native.hpp) and through the generic trampoline (ALPAKA_B200_NATIVE=0);
  * the reference's unmodified ReduceKernel (two launches) and the native single-pass reduction agree with the oracle.
The binaries are built on the CPU box by __graft_entry__.build() and travel with the snapshot; a missing binary FAILS
(there is no fallback)."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import P

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "examples")


def run(exe, *args, native=True, timeout=600, check=True, tune=None):
    path = os.path.join(BIN, exe)
    assert os.path.exists(path), f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'` first"
    env = dict(os.environ)
    env["ALPAKA_B200_NATIVE"] = "1" if native else "0"
    if tune:
        env["B200_TUNE"] = tune
    r = subprocess.run([path, *args], capture_output=True, text=True, timeout=timeout, env=env)
    if check:
        assert r.returncode == 0, f"{exe} {' '.join(args)} -> rc {r.returncode}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r


def last_json(stdout):
    for line in reversed(stdout.splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise AssertionError("no JSON line in output:\n" + stdout[-2000:])


# ---------------------------------------------------------------------------------- unmodified reference drivers
@pytest.mark.parametrize("native", [True, False])
def test_reference_heat_driver_unmodified(native):
    """example/heatEquation2D/src/heatEquation2D.cpp as shipped: 64x64, 4000 steps, max-abs error < 1e-4."""
    r = run("ref_heatEquation2D", native=native)
    assert "TagGpuB200" in r.stdout
    assert "AccGpuB200<2,unsigned int>" in r.stdout
    assert "Execution results correct!" in r.stdout


@pytest.mark.parametrize("native", [True, False])
def test_reference_babelstream_driver_unmodified(native):
    """benchmarks/babelstream/src/babelStreamMainTest.cpp as shipped, float and double, five kernels + Dot
    (the driver's REQUIREs: A=1, B=2, C=5, Dot=2N within 100 eps)."""
    r = run("ref_babelstream", "--array-size=4194304", "--number-runs=5", native=native)
    assert "All tests passed (8 assertions in 2 test cases)" in r.stdout
    # Dot ran: the driver only runs it for the CUDA/HIP/SYCL-GPU tags (babelStreamMainTest.cpp:372)
    assert r.stdout.count("DotKernel") >= 2
    assert "AccGpuB200<1,unsigned int>" in r.stdout


def test_block_hierarchy_atomics_are_atomic_between_blocks():
    """4096 blocks x 256 threads add to ONE global counter under hierarchy::Blocks / Grids (tests/cpp/atomic_blocks.cpp)."""
    r = run("test_atomic_blocks")
    assert "-> OK" in r.stdout, r.stdout


# Reference examples next to the hot path, compiled unmodified (examples/Makefile REFEX):
# each is its own checker (exit status + the success line its main() prints).
REF_EXAMPLES = {
    "heatEquation": "Execution results correct!",          # 1-D FTCS, example/heatEquation/src/heatEquation.cpp:173-177
    "vectorAdd": "Execution results correct!",             # example/vectorAdd/src/vectorAdd.cpp:166-180
    "convolution1D": "All results are correct!",           # example/convolution1D/src/convolution1D.cpp:187
    "convolution2D": "Sampled result checks are correct!", # example/convolution2D/src/convolution2D.cpp:391
    "parallelLoopPatterns": "Test passed.",                # five loop patterns, each checked by testResult() (:34-51)
    "helloWorld": "[z:0, y:0, x:0]",                       # device printf of the first thread
    "helloWorldLambda": None,
    "kernelSpecialization": None,
    "tagSpecialization": None,
    "openMPSchedule": None,
    "ls": "AccGpuB200<1,int>",
}


@pytest.mark.parametrize("name", sorted(REF_EXAMPLES))
def test_reference_example_unmodified(name):
    r = run("ref_ex_" + name)
    marker = REF_EXAMPLES[name]
    out = r.stdout + r.stderr
    assert "incorrect" not in out.lower() and "Test failed" not in out, out[-2000:]
    if marker is not None:
        assert marker in out, out[-2000:]


# ---------------------------------------------------------------------------------- BabelStream parity vs the oracle
@pytest.mark.parametrize("native", [True, False])
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("kernel", ["init", "copy", "mul", "add", "triad", "nstream"])
def test_babelstream_functors_bit_exact(tmp_path, kernel, precision, native):
    dtype = np.float64 if precision == "double" else np.float32
    kind = "uniform_f64" if precision == "double" else "uniform_f32"
    n = (1 << 18) + 1024 * 3
    a, b, c = (ol.fill(kind, n, seed=ol.SEED + 10 + k) for k in range(3))
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([a, b, c]).tofile(inp)
    run("babelstream_b200", f"--parity-kernel={kernel}", f"--precision={precision}", f"--array-size={n}",
        f"--input={inp}", f"--output={out}", native=native)
    got = np.fromfile(out, dtype=dtype)
    ol.orc_stream(kernel, a, b, c, scalar=2.0, init_a=1.0)
    want = np.concatenate([a, b, c])
    assert got[: 3 * n].tobytes() == want.tobytes()


@pytest.mark.parametrize("native", [True, False])
def test_babelstream_ragged_size_bit_exact(tmp_path, native):
    """n with no power-of-two factor beyond 2^0: the work division degenerates, the result must not."""
    n = 1000003 if native else 100003
    a, b, c = (ol.fill("uniform_f64", n, seed=77 + k) for k in range(3))
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([a, b, c]).tofile(inp)
    run("babelstream_b200", "--parity-kernel=triad", f"--array-size={n}", f"--input={inp}", f"--output={out}", native=native)
    got = np.fromfile(out, dtype=np.float64)
    ol.orc_stream("triad", a, b, c)
    assert got[: 3 * n].tobytes() == np.concatenate([a, b, c]).tobytes()


@pytest.mark.parametrize("native", [True, False])
@pytest.mark.parametrize("precision,tol", [("double", 1e-12), ("float", 1e-5)])
def test_babelstream_dot_within_tolerance(tmp_path, precision, tol, native):
    dtype = np.float64 if precision == "double" else np.float32
    kind = "uniform_f64" if precision == "double" else "uniform_f32"
    n = (1 << 20) + 17
    a, b = ol.fill(kind, n, seed=5), ol.fill(kind, n, seed=6)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([a, b, np.zeros(n, dtype=dtype)]).tofile(inp)
    run("babelstream_b200", "--parity-kernel=dot", f"--precision={precision}", f"--array-size={n}", f"--input={inp}",
        f"--output={out}", native=native)
    got = float(np.fromfile(out, dtype=dtype)[3 * n])
    sfx = ol.SFX[np.dtype(dtype)]
    want = float(getattr(ol.oracle(), f"orc_dot_{sfx}")(P(a), P(b), n, 256, 1024, None))
    scale = float(np.sum(np.abs(a.astype(np.float64) * b.astype(np.float64))))
    assert abs(got - want) <= tol * scale


def test_babelstream_benchmark_mode_verifies():
    r = run("babelstream_b200", "--array-size=8388608", "--number-runs=5")
    j = last_json(r.stdout)
    assert j["verified"] is True and j["native"] is True
    assert set(j["gbs"]) == {"init", "copy", "mul", "add", "triad", "dot", "nstream"}


def test_native_and_generic_paths_are_distinct():
    """The A/B switch really switches: the generic trampoline launches b200k::run, which at 2^26 doubles is measurably
    slower than the vectorised native Copy (scalar 8-byte accesses, one element per thread)."""
    on = last_json(run("babelstream_b200", "--array-size=67108864", "--number-runs=10", native=True).stdout)
    off = last_json(run("babelstream_b200", "--array-size=67108864", "--number-runs=10", native=False, tune="generic.coarsen=0").stdout)
    assert on["native"] is True and off["native"] is False
    assert on["gbs"]["copy"] > 1.1 * off["gbs"]["copy"]


# ---------------------------------------------------------------------------------- generic path, unrecognised functors
# build/examples/babelstream_b200_renamed = the same driver with functor names the library does not know (UserCopy, ...):
# every launch takes the generic path. Where the launch site can prove that the pointer arguments lie in distinct
# allocations it picks the block-coarsened, restrict-qualified trampoline (b200k::runCoarse, include/alpaka/b200/Kernel.hpp).
COARSE_N = 1024 * 5003  # 5003 blocks of 1024: above the coarsening threshold, and 5003 % 4 != 0 exercises the tail blocks


@pytest.mark.parametrize("tune", [None, "generic.coarsen=0"], ids=["coarsened", "plain"])
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("kernel", ["init", "copy", "mul", "add", "triad", "nstream"])
def test_unrecognised_functors_bit_exact_on_the_generic_path(tmp_path, kernel, precision, tune):
    dtype = np.float64 if precision == "double" else np.float32
    kind = "uniform_f64" if precision == "double" else "uniform_f32"
    n = COARSE_N
    a, b, c = (ol.fill(kind, n, seed=ol.SEED + 40 + k) for k in range(3))
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([a, b, c]).tofile(inp)
    run("babelstream_b200_renamed", f"--parity-kernel={kernel}", f"--precision={precision}", f"--array-size={n}",
        f"--input={inp}", f"--output={out}", tune=tune)
    got = np.fromfile(out, dtype=dtype)
    ol.orc_stream(kernel, a, b, c, scalar=2.0, init_a=1.0)
    assert got[: 3 * n].tobytes() == np.concatenate([a, b, c]).tobytes()


def test_unrecognised_dot_functor_runs_the_plain_trampoline_and_verifies(tmp_path):
    """UserDot uses shared memory and block synchronisation: never coarsened; the driver's own checks must hold."""
    r = run("babelstream_b200_renamed", "--array-size=8388608", "--number-runs=3")
    j = last_json(r.stdout)
    assert j["verified"] is True
    assert set(j["gbs"]) == {"init", "copy", "mul", "add", "triad", "dot", "nstream"}


def test_unrecognised_functors_reach_85_percent_of_native():
    """a renamed Copy / Triad functor must not fall to the reference's one-element-per-thread speed
    (4.1 / 5.3 TB/s on this GPU). 2^28 doubles per array (2 GiB, far beyond L2)."""
    native = last_json(run("babelstream_b200", "--array-size=268435456", "--number-runs=10").stdout)
    renamed = last_json(run("babelstream_b200_renamed", "--array-size=268435456", "--number-runs=10").stdout)
    plain = last_json(run("babelstream_b200_renamed", "--array-size=268435456", "--number-runs=10", tune="generic.coarsen=0").stdout)
    print({k: (native["gbs"][k], renamed["gbs"][k], plain["gbs"][k]) for k in ("copy", "mul", "add", "triad", "nstream")})
    for k in ("copy", "triad"):
        assert renamed["gbs"][k] >= 0.85 * native["gbs"][k], (k, renamed["gbs"][k], native["gbs"][k])
        assert renamed["gbs"][k] > 1.1 * plain["gbs"][k]


# ---------------------------------------------------------------------------------- reduce
@pytest.mark.parametrize("n", [1, 2, 255, 1000, (1 << 22) + 5])
def test_reduce_u32_reference_kernel_and_native_bit_exact(tmp_path, n):
    x = ol.fill("hash_u32", n, seed=11)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    x.tofile(inp)
    run("reduce_b200", f"--n={n}", "--dtype=u32", "--runs=2", f"--input={inp}", f"--output={out}")
    got = np.fromfile(out, dtype=np.uint32)
    want = np.uint32(int(x.astype(np.uint64).sum()) % 2**32)
    assert got[0] == want, "reference ReduceKernel through the generic trampoline"
    assert got[1] == want, "native single-pass reduction through functor recognition"
    # and the oracle's restatement of the reference's GPU launch shape gives the same number
    assert ol.orc_reduce(x, ol.oracle().orc_reduce_block_count(n, 148, 256), 256, 1) == want


def test_reduce_closed_form_like_the_reference_driver():
    """reduce.cpp:137-148: x[i] = i+1, expected n/2*(n+1) mod 2^32 (the driver checks it itself)."""
    r = run("reduce_b200", "--n=16777216", "--dtype=u32", "--runs=2")
    assert "Results match." in r.stdout


def test_reduce_f32_exactly_representable_sums(tmp_path):
    n = (1 << 22) + 3
    x = ol.fill("bernoulli_f32", n, seed=3)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    x.tofile(inp)
    run("reduce_b200", f"--n={n}", "--dtype=f32", "--runs=2", f"--input={inp}", f"--output={out}")
    got = np.fromfile(out, dtype=np.float32)
    want = float(x.astype(np.float64).sum())
    assert abs(float(got[0]) - want) <= 1e-5 * want and abs(float(got[1]) - want) <= 1e-5 * want


# ---------------------------------------------------------------------------------- heatEquation2D
@pytest.mark.parametrize("mode,native,exact", [("fused", True, True), ("fused2", True, True), ("fused3", True, True), ("fused4", True, True), ("fused6", True, True), ("fused8", True, True),
                                               ("functors", True, True), ("functors", False, False)])
@pytest.mark.parametrize("shape", [(64, 64), (96, 160)])
def test_heat2d_cpp_driver_vs_oracle(tmp_path, mode, native, exact, shape):
    """fused (one and two steps per launch) and recognised-functor paths: bit-exact (boundary factors from the host
    libm). Generic trampoline: the reference's BoundaryKernel calls the DEVICE exp/sin, so only the 1e-12 max-abs bar
    applies. 61 steps: the two-step mode ends with a single-step launch."""
    ny, nx = shape
    steps = 61
    dx, dy, dt = ol.heat_params(ny, nx)
    out = tmp_path / "u.bin"
    r = run("heat2d_b200", f"--ny={ny}", f"--nx={nx}", f"--steps={steps}", f"--dt={dt!r}", f"--mode={mode}", f"--output={out}",
            native=native, check=False)
    assert os.path.exists(out), r.stdout + r.stderr
    got = np.fromfile(out, dtype=np.float64).reshape(ny + 2, nx + 2)
    u0 = np.empty((ny + 2, nx + 2))
    ol.oracle().orc_heat2d_init(P(u0), ny, nx, nx + 2, dx, dy)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    if exact:
        assert got.tobytes() == want.tobytes()
    else:
        assert float(np.max(np.abs(got - want))) <= 1e-12


@pytest.mark.parametrize("levels", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("slabs,shape", [(3, (96, 160)), (2, (256, 700)), (1, (64, 64))])
def test_heat2d_cpp_slabs_vs_oracle(tmp_path, slabs, shape, levels):
    """alpaka::b200::Heat2DSlabs: K row slabs in one process (here all on device 0), `levels` time levels per launch and per
    ghost-row exchange; 62 steps end with shallower launches. Bit-exact against the undecomposed oracle."""
    ny, nx = shape
    steps = 62
    dx, dy, dt = ol.heat_params(ny, nx)
    out = tmp_path / "u.bin"
    r = run("heat2d_b200", f"--ny={ny}", f"--nx={nx}", f"--steps={steps}", f"--dt={dt!r}", "--mode=slabs", f"--slabs={slabs}",
            f"--levels={levels}", f"--output={out}", check=False)
    assert os.path.exists(out), r.stdout + r.stderr
    got = np.fromfile(out, dtype=np.float64).reshape(ny + 2, nx + 2)
    u0 = np.empty((ny + 2, nx + 2))
    ol.oracle().orc_heat2d_init(P(u0), ny, nx, nx + 2, dx, dy)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    assert got.tobytes() == want.tobytes()
    assert last_json(r.stdout)["launches"] == len(__import__("alpaka_b200").decomp.launch_schedule(steps, levels, 2))


@pytest.mark.parametrize("levels", [4, 8])
@pytest.mark.parametrize("grid,shape", [((2, 2), (96, 320)), ((1, 2), (64, 300)), ((3, 1), (120, 140)), ((2, 3), (64, 480))])
def test_heat2d_cpp_tiles_vs_oracle(tmp_path, grid, shape, levels):
    """alpaka::b200::Heat2DTiles: Py x Px tiles in one process (here all on device 0), ghost cells `levels` deep on all sides,
    rows exchanged inside the walker launch and columns + corners by the column kernel. Bit-exact against the undecomposed
    oracle (corners excepted: neither reference kernel writes them)."""
    ny, nx = shape
    steps = 24
    dx, dy, dt = ol.heat_params(ny, nx)
    out = tmp_path / "u.bin"
    r = run("heat2d_b200", f"--ny={ny}", f"--nx={nx}", f"--steps={steps}", f"--dt={dt!r}", "--mode=tiles", f"--py={grid[0]}",
            f"--px={grid[1]}", f"--levels={levels}", f"--output={out}", check=False)
    assert os.path.exists(out), r.stdout + r.stderr
    got = np.fromfile(out, dtype=np.float64).reshape(ny + 2, nx + 2)
    u0 = np.empty((ny + 2, nx + 2))
    ol.oracle().orc_heat2d_init(P(u0), ny, nx, nx + 2, dx, dy)
    want = ol.orc_heat_run(u0, 1, steps, dx, dy, dt)
    assert got.tobytes() == want.tobytes()  # (the driver's host field keeps the initial corners, as the oracle's does)
    assert last_json(r.stdout)["launches"] == 24 // levels


def test_heat2d_cpp_slabs_on_several_devices(tmp_path):
    """The same with one slab per DEVICE of this process (peer stores over NVLink into the other device's pool memory:
    b200_enable_peer_all grants the pools' access). Needs >= 2 devices."""
    import alpaka_b200 as ab

    ndev = ab.Platform().get_dev_count()
    if ndev < 2:
        pytest.skip("needs at least 2 devices")
    slabs = min(ndev, 4)
    ny, nx, steps = 128 * slabs, 900, 30
    dx, dy, dt = ol.heat_params(ny, nx)
    out = tmp_path / "u.bin"
    r = run("heat2d_b200", f"--ny={ny}", f"--nx={nx}", f"--steps={steps}", f"--dt={dt!r}", "--mode=slabs", f"--slabs={slabs}",
            "--levels=3", f"--output={out}", check=False)
    assert os.path.exists(out), r.stdout + r.stderr
    assert last_json(r.stdout)["devices"] == slabs
    got = np.fromfile(out, dtype=np.float64).reshape(ny + 2, nx + 2)
    u0 = np.empty((ny + 2, nx + 2))
    ol.oracle().orc_heat2d_init(P(u0), ny, nx, nx + 2, dx, dy)
    assert got.tobytes() == ol.orc_heat_run(u0, 1, steps, dx, dy, dt).tobytes()


def test_heat2d_reference_configuration_with_run_time_sizes():
    """The shipped configuration (64x64, 4000 steps, tMax 0.1) through the parameterised driver, both modes."""
    for mode in ("functors", "fused", "fused2", "fused3", "fused4"):
        r = run("heat2d_b200", "--ny=64", "--nx=64", "--steps=4000", "--dt=2.5e-05", f"--mode={mode}")
        assert "Execution results correct!" in r.stdout
        assert last_json(r.stdout)["max_error"] < 1e-4
