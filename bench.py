#!/usr/bin/env python
"""bench.py -- BabelStream Triad (headline) + the rest of the hot path on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU implementation (oracle/_ref) on the host cores

Contract (see DESIGN.md "Measurement"):
  * a STEP is one Triad pass c = a + 2*b over double arrays of 2^30 elements per GPU (BASELINE.json configs[1]); the
    arrays (25.8 GB per GPU) are far larger than the 126 MB L2, so no flush is needed between iterations;
  * `value` = whole-job Triad GB/s (24 B/element, all ranks) with inputs resident in HBM, timed with CUDA events on
    the launching stream, barrier + synchronize on both sides, max over ranks;
  * `roofline` = the Triad kernel's algorithmic bytes per launch / its average launch duration vs the measured HBM
    copy bandwidth (MEASURED_PEAKS.json);
  * `e2e` = the same Triad through the public host API with HOST buffers (pinned), H2D and D2H inside the timed
    region, chunk-pipelined over 4 streams; `e2e_paths` = the same for the two drivers that really copy
    (heatEquation2D: upload, 1000 steps, download; reduce: upload, reduce, scalar back);
  * `kernels` = every other kernel of the path (Init/Copy/Mul/Add/Nstream/Dot, reduce u32/f32, heatEquation2D) timed
    the same way, as absolute GB/s and fraction of the HBM roofline, plus the reference's own timing method
    (`gbs_host_min`: host clock around exec + wait, min excluding the first run, babelStreamMainTest.cpp:279-301),
    the reference's CPU back-end on this box's cores (`cpu_gbs`) and, at N = 1, the reference's own CUDA back-end
    (AccGpuCudaRt built for sm_100a, `gpu_reference`) on the same GPU;
  * N > 1: `kernels.*_strong` = BASELINE.json configs[1] as written -- 2^30 elements TOTAL, slab-sharded over the
    ranks -- next to the weak lines; `parity_n` = the sharded paths checked bit for bit against the committed golden
    vectors of the unmodified reference (tests/golden/multi_gpu_vectors.npz) before anything is timed: a mismatch
    aborts the run;
  * `cpu_baseline` = the reference's AccCpuOmp2Blocks Triad (oracle/_ref) on the host cores, bounded sample.
Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "babelstream_triad_gbs"
UNIT = "GB/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
FP64_LANES_PER_SM_CLK = 64  # DFMA/DADD/DMUL issue rate of one B200 SM (ncu: sm__inst_executed_pipe_fp64 peak = 2 warps/clk)
STREAM_BYTES = {"init": 24.0, "copy": 16.0, "mul": 16.0, "add": 24.0, "triad": 24.0, "nstream": 32.0, "dot": 16.0}
REF_KERNEL_ID = {"init": 0, "copy": 1, "mul": 2, "add": 3, "triad": 4, "nstream": 5, "dot": 6}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def bench_config(n: int, world: int) -> dict:
    """The SAME dictionary on both arms (ours and --impl reference)."""
    return {"workload": "BabelStream Triad double, 2^30 elements/array per GPU (BASELINE.json configs[1])",
            "elements_per_gpu": n, "bytes_per_element": 24,
            "l2": "inputs (25.8 GB per GPU) larger than L2, no flush needed",
            "parallelism": f"slab x{world}" if world > 1 else "single GPU"}


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen(
                [exe, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- host / CPU baselines
def _affinity_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# taken once at import: after the reference's OpenMP runtime has started with OMP_PROC_BIND, the calling thread is bound to one
# core and the affinity mask reads 1
_HOST_THREADS = _affinity_threads()


def host_threads():
    return _HOST_THREADS


def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def host_description() -> dict:
    """CPU model, sockets, physical cores and threads visible to this process (the CPU baselines' denominator)."""
    model, sockets, cores = "unknown", set(), set()
    try:
        phys = core = None
        with open("/proc/cpuinfo") as f:
            for line in f:
                k, _, v = line.partition(":")
                k, v = k.strip(), v.strip()
                if k == "model name":
                    model = v
                elif k == "physical id":
                    phys = v
                    sockets.add(v)
                elif k == "core id":
                    core = v
                elif k == "" and phys is not None:
                    cores.add((phys, core))
                    phys = core = None
    except Exception:
        pass
    return {"cpu_model": model, "sockets": len(sockets) or 1, "physical_cores": len(cores) or host_threads(),
            "threads_used": host_threads(),
            "build": "oracle/Makefile: g++ -O3 -fopenmp -ffp-contract=off, no -march=native (FMA contraction pinned off for parity)"}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm must use all the host threads it can. Set the
    environment before the OpenMP runtime of oracle/_ref initialises, and the runtime's ICV in case it already has."""
    nthreads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(nthreads)
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    os.environ.setdefault("OMP_PLACES", "cores")
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(nthreads))
    except OSError:
        pass
    return nthreads


def time_reference_stream(kernel: str, n: int, runs: int):
    """One reference BabelStream kernel functor (Dot: WorkDiv {256,1,1}) on AccCpuOmp2Blocks (oracle/_ref), timed the
    reference's way (host clock around exec + wait, babelStreamMainTest.cpp:279-291). -> seconds per run, threads."""
    import oracle_lib as ol

    L = ol.ref()
    if L is None:
        return None, 0, "oracle/_ref/libalpaka_ref.so missing"
    secs = (C.c_double * runs)()
    rc = L.ref_babelstream_time(REF_KERNEL_ID[kernel], 1, n, runs, secs, 256 if kernel == "dot" else 0)
    if rc != 0:
        return None, 0, f"ref_babelstream_time rc={rc}"
    return list(secs), L.ref_omp_max_threads(), None


def time_port_triad(n: int, runs: int):
    """Fallback CPU baseline: the plain-C oracle port (OpenMP) when oracle/_ref is unavailable."""
    import numpy as np
    import oracle_lib as ol

    L = ol.oracle()
    a, b, c = np.ones(n), np.full(n, 2.0), np.zeros(n)
    out = []
    for _ in range(runs):
        t0 = time.perf_counter()
        L.orc_triad_f64(ol.P(a), ol.P(b), ol.P(c), 2.0, n)
        out.append(time.perf_counter() - t0)
    return out, host_threads()


def min_mean(secs):
    """The reference's statistics: the first run is excluded (babelStreamCommon.hpp:168-206)."""
    t = secs[1:] if len(secs) > 1 else secs
    return min(t), sum(t) / len(t)


def cpu_baselines(quick: bool) -> dict:
    """The reference's CPU back-end (oracle/_ref, AccCpuOmp2Blocks, all host threads) per kernel of the path, bounded
    samples: BabelStream at C1's 2^25 (the reference CPU run of BASELINE.json configs[0]), reduce 2^30 / 2^31 uint32 and
    2^30 float, heat 2048^2 .. 16384^2 a few steps. GB/s by the same byte formulas as the GPU lines."""
    import numpy as np
    import oracle_lib as ol
    from oracle_lib import P

    out = {}
    L = ol.ref()
    if L is None:
        return {"unavailable": "oracle/_ref/libalpaka_ref.so missing"}
    n = 1 << 25
    for k in ("init", "copy", "mul", "add", "triad", "nstream", "dot"):
        secs, _, err = time_reference_stream(k, n, 11)
        if secs is None:
            out[f"{k}_f64"] = {"error": err}
            continue
        mn, mean = min_mean(secs)
        out[f"{k}_f64"] = {"gbs_min": round(STREAM_BYTES[k] * n * 1e-9 / mn, 1), "gbs_mean": round(STREAM_BYTES[k] * n * 1e-9 / mean, 1),
                           "n": n, "runs": 11}
    if quick:
        return out
    sec = C.c_double(0)
    for tag, dtype, sizes in (("u32", np.uint32, (1 << 30, 1 << 31)), ("f32", np.float32, (1 << 30,))):
        for m in sizes:
            if 4.0 * m / 1e9 > 0.3 * host_mem_available_gb():
                continue
            x = np.ones(m, dtype=dtype)
            res = np.zeros(1, dtype=dtype)
            ts = []
            for _ in range(4):
                rc = getattr(L, f"ref_reduce_{tag}")(1, P(x), m, P(res), C.byref(sec))
                ts.append(sec.value)
            mn, mean = min_mean(ts)
            out[f"reduce_{tag}_n2^{m.bit_length() - 1}"] = {"gbs_min": round(4.0 * m * 1e-9 / mn, 1), "gbs_mean": round(4.0 * m * 1e-9 / mean, 1),
                                                             "rc": rc, "result_ok": bool(float(res[0]) == float(m % 2**32 if tag == "u32" else m) or tag == "f32")}
            del x
    for edge, steps in ((2048, 20), (8192, 4), (16384, 2)):
        if 3 * 8.0 * (edge + 2) ** 2 / 1e9 > 0.4 * host_mem_available_gb():
            continue
        dx, dy = 1.0 / (edge + 1), 1.0 / (edge + 1)
        dt = 0.2 * min(dx * dx, dy * dy)
        u = np.empty((edge + 2, edge + 2))
        L.ref_heat2d_init(P(u), edge, edge, dx, dy)
        rc = L.ref_heat2d_run(1, P(u), edge, edge, 1, steps, dx, dy, dt, C.byref(sec))
        out[f"heat2d_{edge}"] = {"gbs": round(16.0 * edge * edge * steps * 1e-9 / sec.value, 1), "ms_per_step": round(sec.value * 1e3 / steps, 3),
                                 "steps": steps, "rc": rc, "note": "Stencil + Boundary functors per step, host clock around the loop + wait"}
        del u
    return out


# ----------------------------------------------------------------------------------------------- reference GPU back-end
def gpu_reference(quick: bool) -> dict:
    """The reference's OWN CUDA back-end (AccGpuCudaRt) built for sm_100a by oracle/Makefile `gpuref` and run on this GPU:
    its babelstream driver as shipped (2^30 doubles) and its heatEquation2D / reduce kernels through oracle/ref_gpu_main.cpp
    (run-time sizes). Separate processes; nothing of this repository's kernels is on that path."""
    d = os.path.join(ROOT, "oracle", "_ref")
    out = {}
    exe = os.path.join(d, "ref_gpu_babelstream")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe, "--array-size=1073741824", "--number-runs=10", "TEST: Babelstream Five Kernels<Double>*"],
                               capture_output=True, text=True, timeout=600)
            rows = {}
            for line in r.stdout.splitlines():
                parts = line.split()
                if len(parts) >= 6 and parts[0].endswith("Kernel"):
                    rows[parts[0]] = {"gbs_min": float(parts[1]), "min_s": float(parts[2]), "avg_s": float(parts[4])}
            names = {"InitKernel": "init", "CopyKernel": "copy", "MultKernel": "mul", "AddKernel": "add", "TriadKernel": "triad", "DotKernel": "dot"}
            for k, v in rows.items():
                key = names.get(k)
                if key is None:
                    continue
                gbs_mean = v["gbs_min"] * v["min_s"] / v["avg_s"] if v["avg_s"] > 0 else None
                out[f"{key}_f64"] = {"gbs_min": v["gbs_min"], "gbs_mean": round(gbs_mean, 1) if gbs_mean else None}
            if "init_f64" in out:
                out["init_f64"]["note"] = "the reference books 2 arrays for Init (babelStreamMainTest.cpp:415); x1.5 for the 24 B/element of the GPU line"
            out["babelstream_rc"] = r.returncode
            if r.returncode != 0:
                out["babelstream_note"] = ("the driver's own A/B/C check sums 2^30 errors in the element type and trips at this size; "
                                           "the timings above are printed regardless")
        except Exception as e:  # noqa: BLE001
            out["babelstream_error"] = str(e)[:200]
    else:
        out["babelstream_error"] = "oracle/_ref/ref_gpu_babelstream missing (built by `make -C oracle gpuref` where /root/reference exists)"
    exe = os.path.join(d, "ref_gpu")
    if os.path.exists(exe) and not quick:
        for name, args in (("heat2d_f64", ["heat", "16384", "16384", "100"]), ("reduce_u32", ["reduce_u32", str(1 << 32), "6"]),
                           ("reduce_f32", ["reduce_f32", str(1 << 30), "6"])):
            try:
                r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                out[name] = json.loads(line[-1]) if line else {"error": (r.stderr or r.stdout)[-200:]}
            except Exception as e:  # noqa: BLE001
                out[name] = {"error": str(e)[:200]}
    return out


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    use_all_host_threads()
    n_full = 1 << 30
    n = args.n or n_full
    # 3 arrays of n doubles must fit the host comfortably
    while 3 * 8 * n / 1e9 > 0.5 * host_mem_available_gb() and n > (1 << 22):
        n //= 2
    runs = args.warmup + args.steps
    kind = "reference"
    secs, threads, err = time_reference_stream("triad", n, runs)
    if secs is None:
        kind = "port"
        secs, threads = time_port_triad(n, runs)
    timed = secs[args.warmup:]
    mean_s = sum(timed) / len(timed)
    value = 24.0 * n * 1e-9 / mean_s
    sample = (f"reference TriadKernel (babelStreamMainTest.cpp:125-141) on AccCpuOmp2Blocks<1,uint32>, double, "
              f"n=2^{n.bit_length() - 1} per step, {threads} OpenMP threads, host clock around exec+wait; value = mean over the "
              f"timed steps, best step {24.0 * n * 1e-9 / min(timed):.1f} GB/s")
    if kind == "port":
        sample = f"oracle C port (OpenMP) because: {err}; n={n}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(n_full, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "value_min_method": 24.0 * n * 1e-9 / min(timed), "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample, "elements_per_step": n, "host": host_description()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import math

    import numpy as np

    import alpaka_b200 as ab
    from alpaka_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    platform = ab.Platform()
    if platform.get_dev_count() <= local_rank:
        raise RuntimeError("bench.py needs one CUDA device per rank; alpaka_b200 has no CPU fallback")
    dev = platform.get_dev_by_idx(local_rank)
    q = ab.Queue(dev)
    _lib.load()
    peak, peak_kind = measured_peak()
    t_start = time.perf_counter()

    def barrier():
        q.wait()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(x: float) -> list:
        if dist is None:
            return [x]
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(p.item()) for p in parts]

    e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)

    def timed(fn, steps, warmup):
        """W untimed + exactly K timed launches between barriers; CUDA events on the launching stream."""
        for _ in range(warmup):
            fn()
        barrier()
        ab.enqueue(q, e0)
        for _ in range(steps):
            fn()
        ab.enqueue(q, e1)
        q.wait()
        ms = e0.elapsed_ms(e1)
        barrier()
        return max_over_ranks(ms) / steps

    def host_min(fn, runs=11):
        """The reference's own timing method (measureKernelExec, babelStreamMainTest.cpp:279-301): host clock around
        exec + wait per run, minimum excluding the first run. N = 1 only (a per-rank wall clock is not a job time)."""
        ts = []
        for _ in range(runs):
            t0 = time.perf_counter()
            fn()
            q.wait()
            ts.append(time.perf_counter() - t0)
        return min(ts[1:])

    def timed_run(fn, units):
        """ONE long call (e.g. the 1000 steps BASELINE.json's heat configs specify) between barriers, CUDA events on the
        launching stream, clocks sampled meanwhile: the SUSTAINED figure next to the short-burst ones."""
        smp = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            smp.start()
        ab.enqueue(q, e0)
        fn()
        ab.enqueue(q, e1)
        q.wait()
        ms = e0.elapsed_ms(e1)
        clk = smp.stop() if rank == 0 else None
        barrier()
        return max_over_ranks(ms) / units, clk

    # ---- N > 1: parity of every sharded path against the committed golden vectors of the unmodified reference, BEFORE
    # anything is timed. A mismatch raises and the run fails (no line is printed).
    parity_n = None
    if world > 1 and not args.no_parity:
        import golden_multi

        parity_n = golden_multi.check_all(golden_multi.Ranks(ab, world, {rank: q}, dist))
        barrier()

    n = args.n or (1 << 30)
    K, W = args.steps, max(args.warmup, 3)
    bs = ab.babelstream

    a, b, c = (ab.alloc_buf(dev, np.float64, n, q) for _ in range(3))
    # the driver's own data: a = 1, then b = 2 (Copy + Mult), c = 5 after Triad (babelStreamMainTest.cpp:305-339)
    bs.init(q, a, b, c)
    bs.copy(q, a, b)
    bs.mul(q, a, b)
    q.wait()

    # ---- headline: Triad
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ab.runtime.launch_count()
    ms_triad = timed(lambda: bs.triad(q, a, b, c), K, W)
    launches = ab.runtime.launch_count() - launches0 - W
    clocks = sampler.stop() if rank == 0 else None
    triad_bytes = 24.0 * n
    value = world * triad_bytes * 1e-9 / (ms_triad * 1e-3)

    # spot-check the result the timed kernel produced (c must be 1 + 2*2 = 5 everywhere in a sampled window)
    chk = np.empty(1 << 16)
    ab.memcpy(q, chk, ab.create_view(dev, c.ptr + 8 * (n - (1 << 16)), np.float64, 1 << 16))
    q.wait()
    assert (chk == 5.0).all(), "Triad result check failed"

    # ---- every other kernel of the path, same timing method (fewer steps for the long ones)
    kernels = {}

    def record(name, ms, nbytes, per_gpu_bytes=None):
        gbs = world * nbytes * 1e-9 / (ms * 1e-3)
        kernels[name] = {"gbs": round(gbs, 1), "ms": round(ms, 4), "frac_of_hbm_peak": round(gbs / world / peak, 4)}
        return kernels[name]

    record("triad_f64", ms_triad, triad_bytes)
    exch = None
    if not args.quick and not args.only_heat:
        Ks = max(5, K // 2)
        stream_calls = {
            "init": lambda: bs.init(q, a, b, c), "copy": lambda: bs.copy(q, a, c), "mul": lambda: bs.mul(q, a, b),
            "add": lambda: bs.add(q, a, b, c), "nstream": lambda: bs.nstream(q, c, a, b, 0.0),
        }
        record("init_f64", timed(stream_calls["init"], Ks, 3), 24.0 * n)
        bs.copy(q, a, b)
        bs.mul(q, a, b)
        record("copy_f64", timed(stream_calls["copy"], Ks, 3), 16.0 * n)
        record("mul_f64", timed(stream_calls["mul"], Ks, 3), 16.0 * n)
        record("add_f64", timed(stream_calls["add"], Ks, 3), 24.0 * n)
        record("nstream_f64", timed(stream_calls["nstream"], Ks, 3), 32.0 * n)
        out = ab.alloc_buf(dev, np.float64, 1, q)
        exch = None
        if dist is None:
            record("dot_f64", timed(lambda: bs.dot_async(q, a, b, out), Ks, 3), 16.0 * n)
            # the reference's method beside the CUDA-event means
            stream_calls["triad"] = lambda: bs.triad(q, a, b, c)
            stream_calls["dot"] = lambda: bs.dot_async(q, a, b, out)
            for k, fn in stream_calls.items():
                kernels[f"{k}_f64"]["gbs_host_min"] = round(STREAM_BYTES[k] * n * 1e-9 / host_min(fn), 1)
            # the Triad figure does not depend on the data: the same launch on hashed U[-1,1) bits (SURVEY.md 8d parity mode)
            rng = np.random.default_rng(0x5EED)
            blk = rng.uniform(-1.0, 1.0, 1 << 22)
            for buf in (a, b):
                for off in range(0, n, 1 << 22):
                    ab.memcpy(q, ab.create_view(dev, buf.ptr + 8 * off, np.float64, min(1 << 22, n - off)), blk[: min(1 << 22, n - off)])
                blk = blk[::-1].copy()
            q.wait()
            kernels["triad_f64"]["gbs_random_data"] = round(24.0 * n * 1e-6 / timed(lambda: bs.triad(q, a, b, c), Ks, 3), 1)
            bs.init(q, a, b, c)
            bs.copy(q, a, b)
            bs.mul(q, a, b)
        else:
            # Dot's exchange step (one scalar per GPU, combined in rank order) is FUSED into the reduction launch: peer
            # stores + flag words from the last block, no NCCL and no host step inside the timed region
            from alpaka_b200 import multi as _multi

            exch = _multi.ScalarExchange(q, rank, world)
            _multi.connect_exchange_over_process_group(exch, dist)
            record("dot_f64", timed(lambda: exch.dot_async(q, a, b, out), Ks, 3), 16.0 * n)
            kernels["dot_f64"]["exchange"] = "all ranks, fused into the launch (peer stores + flags), rank-ordered"
        dot_host = np.empty(1)
        ab.memcpy(q, dot_host, out)
        q.wait()
        dot_total = float(dot_host[0])
        assert dot_total == 2.0 * n * world, f"Dot check failed: {dot_total} != {2.0 * n * world}"

        if dist is not None:
            # ---- BASELINE.json configs[1] as written: 2^30 elements TOTAL, slab-sharded over the ranks (strong scaling).
            # Per-GPU kernels shrink to 1/N (0.45 ms at N = 8): launch ramp and tail are inside the figure.
            from alpaka_b200 import decomp as _decomp

            n_total = n
            lo, hi = _decomp.slab_bounds(n_total, world, rank, align=4)
            m = hi - lo
            Kt = max(20, K)
            for name, fn, bpe in (("triad_f64_strong", lambda: bs.triad(q, a, b, c, n=m), 24.0),
                                  ("copy_f64_strong", lambda: bs.copy(q, a, c, n=m), 16.0),
                                  ("dot_f64_strong", lambda: exch.dot_async(q, a, b, out, n=m), 16.0)):
                ms = timed(fn, Kt, 5)
                gbs = bpe * n_total * 1e-9 / (ms * 1e-3)
                kernels[name] = {"gbs": round(gbs, 1), "ms": round(ms, 4), "frac_of_hbm_peak": round(gbs / world / peak, 4),
                                 "scaling": "strong", "elements_total": n_total, "elements_per_gpu": m}
            ab.memcpy(q, dot_host, out)
            q.wait()
            assert float(dot_host[0]) == 2.0 * n_total, f"strong-scaled Dot check failed: {float(dot_host[0])} != {2.0 * n_total}"
            kernels["dot_f64_strong"]["exchange"] = kernels["dot_f64"]["exchange"]
        out.free()

    for buf in (a, b, c):
        buf.free()
    q.wait()

    e2e_paths = {}
    if not args.quick:
        if not args.only_heat:
            # ---- example/reduce: 2^32 uint32 (17.2 GB) and 2^30 float (BASELINE.json configs[2])
            nr = (1 << 32) if not args.n else 4 * n
            src = ab.alloc_buf(dev, np.uint32, nr, q)
            res = ab.alloc_buf(dev, np.uint32, 1, q)
            ab.memset(q, src, 1)  # every element 0x01010101
            Ks = max(5, K // 2)
            if exch is None:
                record("reduce_u32", timed(lambda: ab.reduce.reduce_sum_async(q, src, res), Ks, 3), 4.0 * nr)
                kernels["reduce_u32"]["gbs_host_min"] = round(4.0 * nr * 1e-9 / host_min(lambda: ab.reduce.reduce_sum_async(q, src, res)), 1)
            else:
                record("reduce_u32", timed(lambda: exch.reduce_sum_async(q, src, res), Ks, 3), 4.0 * nr)
                kernels["reduce_u32"]["exchange"] = kernels["dot_f64"]["exchange"]
            got = np.empty(1, dtype=np.uint32)
            ab.memcpy(q, got, res)
            q.wait()
            assert int(got[0]) == (0x01010101 * nr * world) % 2**32, "reduce u32 check failed"

            if world == 1 and not args.no_e2e:
                # e2e of the reduce driver (reduce.cpp:137-147 uploads, then reduces, then reads the scalar): pinned host
                # source, H2D in 4 slices on the queue, one single-pass reduction, 4 bytes back; wall clock
                ne = nr
                while 4.0 * ne / 1e9 > 0.3 * host_mem_available_gb() and ne > (1 << 26):
                    ne //= 2
                hsrc = ab.alloc_mapped_buf(np.uint32, ne)
                hsrc.array[:] = 1
                lib = _lib.load()

                def reduce_e2e():
                    sl = ne // 4
                    for k4 in range(4):
                        lib.b200_memcpy_async(dev.idx, src.ptr + 4 * k4 * sl, hsrc.ptr + 4 * k4 * sl, 4 * sl, 1, q.handle)
                    ab.reduce.reduce_sum_async(q, src, res, n=ne)
                    ab.memcpy(q, got, res)
                    q.wait()

                reduce_e2e()
                t0 = time.perf_counter()
                for _ in range(2):
                    reduce_e2e()
                t_e = (time.perf_counter() - t0) / 2
                assert int(got[0]) == ne % 2**32, "e2e reduce check failed"
                e2e_paths["reduce_u32"] = {"gbs": round(4.0 * ne * 1e-9 / t_e, 1), "ms": round(t_e * 1e3, 2), "elements": ne,
                                           "h2d_bytes": 4 * ne, "d2h_bytes": 4,
                                           "how": "pinned host source -> H2D -> single-pass reduction -> scalar back, wall clock (reduce.cpp:137-147)"}
                del hsrc
            src.free()
            res.free()
            nf = (1 << 30) if not args.n else n
            srcf = ab.alloc_buf(dev, np.float32, nf, q)
            resf = ab.alloc_buf(dev, np.float32, 1, q)
            ab.memset(q, srcf, 0)
            if exch is None:
                record("reduce_f32", timed(lambda: ab.reduce.reduce_sum_async(q, srcf, resf), Ks, 3), 4.0 * nf)
            else:
                record("reduce_f32", timed(lambda: exch.reduce_sum_async(q, srcf, resf), Ks, 3), 4.0 * nf)
                kernels["reduce_f32"]["exchange"] = kernels["dot_f64"]["exchange"]
                assert exch.status() == 0, "a rank's flag never arrived in the fused Dot/reduce exchange"
                barrier()
                exch.close()
            for bf in (srcf, resf):
                bf.free()

        # ---- heatEquation2D 16384^2 double (BASELINE.json configs[3]); 16 B per core cell per step.
        # N = 1: the plain fused step. N > 1: STRONG scaling of the same 16384^2 field, Py x Px decomposition, halo
        # exchange fused into the step kernel over CUDA-IPC peer pointers (no NCCL, no host synchronisation per step).
        NY = NX = args.heat or 16384

        def fp64_view(entry, levels, cells, ms_launch, clk):
            """The fused launches are bound by the FP64 pipe and the issue slots next to HBM: algorithmic DP instructions (on
            square cells 2 mul + 4 add per cell and level, every product shared by the cells that use it) over the pipe's
            issue rate at the clock seen."""
            mhz = (clk or {}).get("sm_mhz") or 1965.0
            rate = 148 * FP64_LANES_PER_SM_CLK * mhz * 1e6
            entry["fp64_pipe_frac"] = round(6.0 * cells * levels / (ms_launch * 1e-3) / rate, 4)
            entry["fp64_pipe_note"] = (f"6 DP instr per cell and level (algorithmic minimum, square cells) / (148 SMs x 64 lanes x {mhz:.0f} MHz); "
                                       "executed: x 128/120 (4 levels; 128/112 at 6, 8) columns a warp window gives up, x (rows + 2 levels)/rows per "
                                       "walker segment; the tile kernels (2, 3 levels, heat.walk=0) about 1.45 x (profiles/r02)")

        if world == 1:
            dx, dy = 1.0 / (NX + 1), 1.0 / (NY + 1)
            dt = 0.2 * min(dx * dx, dy * dy)
            h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
            # the reference driver's initial condition (analyticalSolution.hpp:58-71), NOT zeros: the power an FP64 kernel
            # draws, and with it the clocks of a long run, depend on the operand bits
            hfield = ab.alloc_mapped_buf(np.float64, (NY + 2, NX + 2))
            hfield.array[:] = ab.heat2d.initial_field(NY, NX, dx, dy)
            h.upload(hfield)
            record("heat2d_f64", timed(lambda: h.step(1), max(20, K), 5), 16.0 * NY * NX)
            # 2, 3 and 4 time levels per launch (b200_heat2d_step2_f64 / b200_heat2d_stepn_f64): the same 16 B per cell per
            # step of ALGORITHMIC bytes, a half / a third / a quarter of them actually moved, so the fraction of the HBM peak exceeds 1
            D = h.default_depth()  # 8 at 16384^2 (walker kernel), see alpaka_b200/heat2d.py
            for G in (8, 6, 4, 3, 2):  # walker kernel (8, 6, 4), tile kernels (3, 2)
                kG = record(f"heat2d_f64_{G}_steps_per_launch", timed(lambda: h.step(G, fuse=G), max(20, K), 5), G * 16.0 * NY * NX)
                kG["ms_per_step"] = round(kG["ms"] / G, 4)
                kG["kernel"] = "heatWalkKernel" if G in (4, 6, 8) else ("heatStepNKernel" if G == 3 else "heatStep2Kernel")
                # what one launch actually moves: one read + one write of the field (ncu: 4.25 GB at 16384^2), whatever G
                kG["hbm_pass_gbs"] = round(16.0 * NY * NX * 1e-9 / (kG["ms"] * 1e-3), 1)
                kG["hbm_pass_frac_of_peak"] = round(kG["hbm_pass_gbs"] / peak, 4)
                fp64_view(kG, G, NY * NX, kG["ms"], None)
            # the round-1 tile kernel at four levels, for the record (heat.walk = 0)
            ab.runtime.tune_set("heat.walk", 0)
            kT = record("heat2d_f64_4_steps_per_launch_tile_kernel", timed(lambda: h.step(4, fuse=4), max(20, K), 5), 4 * 16.0 * NY * NX)
            kT.update(ms_per_step=round(kT["ms"] / 4, 4), kernel="heatStepNKernel<4>")
            ab.runtime.tune_set("heat.walk", 1)
            # BASELINE.json configs[3] as specified: 1000 FTCS steps in one go (sustained clocks, not a short burst)
            for G in (() if args.no_sustained else (1, D)):
                ms_step, clk = timed_run(lambda: h.step(1000, fuse=G), 1000)
                kS = record(f"heat2d_f64_1000_steps_{G}_per_launch", ms_step, 16.0 * NY * NX)
                kS.update(sustained=True, ms_per_step=round(ms_step, 4), clocks=clk)
                if G > 1:
                    fp64_view(kS, G, NY * NX, ms_step * G, clk)
            if not args.no_e2e and not args.no_sustained:
                # e2e of the heat driver (heatEquation2D.cpp:88-190): pinned host field up, 1000 steps, field down; wall clock
                def heat_e2e():
                    h.step_index = 0
                    h.upload(hfield)
                    h.step(1000)
                    return h.download(hfield)

                t0 = time.perf_counter()
                final = heat_e2e()
                t_e = time.perf_counter() - t0
                err = ab.heat2d.validate_solution(final, dx, dy, 1000 * dt)
                assert err < 1e-4, f"e2e heat field deviates from the analytic solution: {err}"
                fb = 8 * (NY + 2) * (NX + 2)
                e2e_paths["heat2d_f64_1000_steps"] = {
                    "gbs": round(16.0 * NY * NX * 1000 * 1e-9 / t_e, 1), "seconds": round(t_e, 4), "h2d_bytes": fb, "d2h_bytes": fb,
                    "max_abs_error_vs_analytic": err,
                    "how": f"pinned host field -> H2D -> 1000 FTCS steps ({D} levels per launch) -> D2H, wall clock (heatEquation2D.cpp:88-190)"}
            h.close()
            del hfield
        else:
            from alpaka_b200 import decomp, multi

            tile = decomp.tile_for(rank, world, NY, NX)
            runner = multi.HeatTile(q, tile, NY, NX)
            multi.connect_over_process_group(runner, dist)
            runner.upload(runner.initial_field())
            barrier()
            ms_heat = timed(lambda: runner.step(1), max(50, K), 5)
            assert runner.status() == 0, "heat halo flag wait timed out"
            record("heat2d_f64", ms_heat, 16.0 * NY * NX / world)
            kernels["heat2d_f64"]["scaling"] = "strong"
            kernels["heat2d_f64"]["decomposition"] = f"{tile.py}x{tile.px} tiles of {tile.ny}x{tile.nx}, fused P2P halo"
            # spot check: the decomposed field still satisfies the analytic solution to the reference's tolerance
            local = runner.download()
            tmax = runner.h.step_index * runner.h.dt

            sx, sy = ab.heat2d.boundary_tables(tile.ny, tile.nx, runner.dx, runner.dy, tile.j_offset, tile.i_offset)
            exact = math.exp(-math.pi * math.pi * tmax) * (sx[None, :] + sy[:, None])
            err = float(np.max(np.abs(local[1:-1, 1:-1] - exact[1:-1, 1:-1])))
            assert err < 1e-4, f"decomposed heat field deviates from the analytic solution: {err}"
            kernels["heat2d_f64"]["max_abs_error_vs_analytic"] = err
            runner.close()
            # the same field as row slabs advanced G time levels per launch and per exchange (ghost rows G deep):
            # algorithmic bytes stay 16 B per cell per step, 1/G of them are moved
            slab = multi.HeatSlab(q, rank, world, NY, NX)
            multi.connect_over_process_group(slab, dist)
            slab.upload(slab.initial_field())
            barrier()
            G = slab.levels
            ms_slab = timed(lambda: slab.step(G), max(50, K), 5)
            assert slab.status() == 0, "heat slab flag wait timed out"
            name_s = f"heat2d_f64_{G}_steps_per_launch"
            kS = record(name_s, ms_slab, G * 16.0 * NY * NX / world)
            kS.update(scaling="strong", ms_per_step=round(ms_slab / G, 4),
                      decomposition=f"{world} row slabs of {NY // world}x{NX}, ghost rows {G} deep, fused P2P halo",
                      halo_bytes_per_launch_per_rank=2 * G * (NX + 2) * 8 * (2 if 0 < rank < world - 1 else 1))
            fp64_view(kS, G, NY * NX / world, ms_slab, None)
            if not args.no_sustained:
                ms_step, clk = timed_run(lambda: slab.step(1000), 1000)  # C4 as specified: 1000 steps in launches of slab.levels
                assert slab.status() == 0, "heat slab flag wait timed out"
                kL = record(f"heat2d_f64_1000_steps_{G}_per_launch", ms_step, 16.0 * NY * NX / world)
                kL.update(scaling="strong", sustained=True, ms_per_step=round(ms_step, 4), clocks=clk)
                fp64_view(kL, G, NY * NX / world, ms_step * G, clk)
            local = slab.download()
            tmax = slab.step_index * slab.dt
            exact = math.exp(-math.pi * math.pi * tmax) * (slab.sx[None, :] + slab.sy[:, None])
            err = float(np.max(np.abs(local[G:-G, 1:-1] - exact[G:-G, 1:-1])))
            assert err < 1e-4, f"slab-decomposed heat field deviates from the analytic solution: {err}"
            kernels[name_s]["max_abs_error_vs_analytic"] = err
            slab.close()
            # the same field as 2-D TILES advanced G levels per launch (ghost cells G deep on all sides; rows exchanged inside
            # the walker launch, columns and corners by the column kernel that follows it): north_star's 2-D decomposition
            # with temporal blocking
            def bench_deep_tiles():
                deep = multi.HeatTileDeep(q, rank, world, NY, NX, levels=G)
                multi.connect_over_process_group(deep, dist)
                deep.upload(deep.initial_field())
                barrier()
                ms_deep = timed(lambda: deep.step(G), max(50, K), 5)
                assert deep.status() == 0, "deep tile flag wait timed out"
                name_d = f"heat2d_f64_tiles_{G}_steps_per_launch"
                kD = record(name_d, ms_deep, G * 16.0 * NY * NX / world)
                kD.update(scaling="strong", ms_per_step=round(ms_deep / G, 4),
                          decomposition=f"{deep.tile.py}x{deep.tile.px} tiles of {deep.tile.ny}x{deep.tile.nx}, ghost cells {G} deep, "
                                        "rows fused into the launch + column kernel")
                if not args.no_sustained:
                    ms_step, clk = timed_run(lambda: deep.step(1000), 1000)
                    assert deep.status() == 0, "deep tile flag wait timed out"
                    kDL = record(f"heat2d_f64_tiles_1000_steps_{G}_per_launch", ms_step, 16.0 * NY * NX / world)
                    kDL.update(scaling="strong", sustained=True, ms_per_step=round(ms_step, 4), clocks=clk)
                local = deep.download()
                tmax = deep.step_index * deep.dt
                exact = math.exp(-math.pi * math.pi * tmax) * (deep.sx[None, :] + deep.sy[:, None])
                err = float(np.max(np.abs(local[G:-G, G:-G] - exact[G:-G, G:-G])))
                assert err < 1e-4, f"tile-decomposed heat field (deep ghosts) deviates from the analytic solution: {err}"
                kernels[name_d]["max_abs_error_vs_analytic"] = err
                deep.close()

            try:
                bench_deep_tiles()
            except ab.B200Error as e:  # (--heat with tiles smaller than two ghost depths: the same refusal on every rank)
                kernels["heat2d_f64_tiles"] = {"skipped": str(e)}
            if not args.heat:
                # BASELINE.json configs[4]: 65536^2 weak-scaled over 8 GPUs = 16384 x 32768 core cells per GPU; the same
                # per-GPU tile at other N (halo/interior overlap stress: 2 x 4.3 GB of state per GPU)
                py, px = decomp.process_grid(world)
                NYw, NXw = 16384 * py, 32768 * px
                tile_w = decomp.tile_for(rank, world, NYw, NXw)
                rw = multi.HeatTile(q, tile_w, NYw, NXw)
                multi.connect_over_process_group(rw, dist)
                rw.upload(rw.initial_field())
                barrier()
                ms_w = timed(lambda: rw.step(1), max(20, K), 5)
                assert rw.status() == 0, "heat halo flag wait timed out"
                record("heat2d_f64_weak", ms_w, 16.0 * NYw * NXw / world)
                kernels["heat2d_f64_weak"]["scaling"] = "weak"
                kernels["heat2d_f64_weak"]["decomposition"] = f"{NYw}x{NXw} global, {py}x{px} tiles of {tile_w.ny}x{tile_w.nx}"
                rw.close()
                sw = multi.HeatSlab(q, rank, world, NYw, NXw)
                multi.connect_over_process_group(sw, dist)
                sw.upload(sw.initial_field())
                barrier()
                ms_sw = timed(lambda: sw.step(G), max(20, K), 5)
                assert sw.status() == 0, "heat slab flag wait timed out"
                kW = record(f"heat2d_f64_weak_{G}_steps_per_launch", ms_sw, G * 16.0 * NYw * NXw / world)
                kW.update(scaling="weak", ms_per_step=round(ms_sw / G, 4),
                          decomposition=f"{NYw}x{NXw} global, {world} row slabs of {NYw // world}x{NXw}, ghost rows {G} deep")
                fp64_view(kW, G, NYw * NXw / world, ms_sw, None)
                sw.close()
        q.wait()

    # ---- e2e: Triad through the public host API, pinned HOST buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        ne = args.e2e_n or n
        while 3 * 8 * ne / 1e9 > 0.4 * host_mem_available_gb() / max(world, 1) and ne > (1 << 24):
            ne //= 2
        ha, hb, hc = (ab.alloc_mapped_buf(np.float64, ne) for _ in range(3))
        ha.array[:] = 1.0
        hb.array[:] = 2.0
        pipe = bs.TriadHostPipeline(dev, np.float64)
        pipe.run(ha.array, hb.array, hc.array)  # warm-up
        barrier()
        steps_e = max(1, min(K, 3))
        t0 = time.perf_counter()
        for _ in range(steps_e):
            h2d, d2h = pipe.run(ha.array, hb.array, hc.array)
        t_mine = (time.perf_counter() - t0) / steps_e
        per_rank = gather_floats(t_mine)
        t_e = max(per_rank)
        assert float(hc.array[0]) == 5.0 and float(hc.array[-1]) == 5.0
        e2e = {"value": world * 24.0 * ne * 1e-9 / t_e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "elements_per_step": ne, "steps": steps_e, "ms_per_step": t_e * 1e3,
               "h2d_gbs_per_rank": [round(16.0 * ne * 1e-9 / t, 1) for t in per_rank],
               "bound": "PCIe Gen5 x16 host->device (16 of the 24 algorithmic bytes per element cross it; 51.5 GB/s measured per GPU "
                        "alone, profiles/r01/e2e_probe.log); at N > 1 the ranks share one virtual NUMA node's host memory",
               "how": "pinned host arrays -> 4-stream chunk pipeline (H2D a,b; Triad; D2H c), wall clock incl. sync"}
        pipe.close()
        del ha, hb, hc

    # ---- CPU baselines beside it (rank 0, N=1 only): the reference's AccCpuOmp2Blocks functors, bounded samples
    cpu = None
    host = None
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu:
        use_all_host_threads()
        host = host_description()
        n_cpu, runs = 1 << 25, 21
        secs, threads, err = time_reference_stream("triad", n_cpu, runs)
        kind = "reference"
        if secs is None:
            kind = "port"
            secs, threads = time_port_triad(n_cpu, runs)
        best, mean = min_mean(secs)
        cpu = {"value": 24.0 * n_cpu * 1e-9 / best, "value_mean": 24.0 * n_cpu * 1e-9 / mean, "unit": UNIT, "cores": threads, "kind": kind,
               "host": host,
               "sample": f"Triad double n=2^25 (BASELINE.json configs[0]), {runs} runs, value = min excluding the first "
                         f"(reference method), value_mean = mean of the same runs, AccCpuOmp2Blocks" + ("" if kind == "reference" else f" [port: {err}]")}
        per_kernel = cpu_baselines(args.quick)
        for name, v in per_kernel.items():
            tgt = kernels.get(name)
            if tgt is not None and "gbs_min" in v:
                tgt["cpu_gbs"] = v["gbs_min"]
                tgt["cpu_gbs_mean"] = v["gbs_mean"]
        for name in ("reduce_u32_n2^30", "reduce_u32_n2^31", "reduce_f32_n2^30", "heat2d_2048", "heat2d_8192", "heat2d_16384"):
            if name in per_kernel:
                cpu.setdefault("other_paths", {})[name] = per_kernel[name]
        if "reduce_u32" in kernels and "reduce_u32_n2^31" in per_kernel:
            kernels["reduce_u32"]["cpu_gbs"] = per_kernel["reduce_u32_n2^31"]["gbs_min"]
        if "reduce_f32" in kernels and "reduce_f32_n2^30" in per_kernel:
            kernels["reduce_f32"]["cpu_gbs"] = per_kernel["reduce_f32_n2^30"]["gbs_min"]
        if "heat2d_16384" in per_kernel:
            for name, v in kernels.items():
                if name.startswith("heat2d"):
                    v["cpu_gbs"] = per_kernel["heat2d_16384"]["gbs"]
    if rank == 0 and world == 1 and not args.no_gpu_ref:
        q.wait()
        gpu_ref = gpu_reference(args.quick)
        for name, v in gpu_ref.items():
            tgt = kernels.get(name)
            if tgt is None or not isinstance(v, dict):
                continue
            ref_gbs = v.get("gbs_min", v.get("gbs"))
            if ref_gbs:
                if name == "init_f64":
                    ref_gbs *= 1.5
                tgt["gpu_reference_gbs"] = round(ref_gbs, 1)
                tgt["vs_gpu_reference"] = round(tgt.get("gbs_host_min", tgt["gbs"]) / ref_gbs, 3)
        if "heat2d_f64" in gpu_ref and "gbs" in gpu_ref["heat2d_f64"]:
            for name, v in kernels.items():
                if name.startswith("heat2d_f64_"):
                    v["gpu_reference_gbs"] = gpu_ref["heat2d_f64"]["gbs"]
                    v["vs_gpu_reference"] = round(v["gbs"] / gpu_ref["heat2d_f64"]["gbs"], 3)

    if rank == 0:
        traffic = None
        prof = os.path.join(ROOT, "profiles", "triad_traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        achieved = triad_bytes * 1e-9 / (ms_triad * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_triad, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(n, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_kind": f"{peak_kind} (burst copy figure; kernel timed alone)",
                         "kernel": "streamKernel<TriadOp<double>>"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "kernels": kernels,
        }
        if e2e_paths:
            line["e2e_paths"] = e2e_paths
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        if parity_n is not None:
            line["parity_n"] = parity_n
        line["bench_wall_s"] = round(time.perf_counter() - t_start, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="elements per array per GPU (default 2^30)")
    ap.add_argument("--heat", type=int, default=0, help="heat2d grid edge (default 16384)")
    ap.add_argument("--e2e-n", type=int, default=0)
    ap.add_argument("--quick", action="store_true", help="Triad only")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gpu-ref", action="store_true", help="skip the reference's own CUDA back-end (oracle/_ref/ref_gpu*)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the golden-vector parity block (profiling only)")
    ap.add_argument("--only-heat", action="store_true", help="Triad headline + the heatEquation2D lines only")
    ap.add_argument("--no-sustained", action="store_true", help="skip the 1000-step heat runs (profiling under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
