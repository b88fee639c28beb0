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
    region, chunk-pipelined over 4 streams;
  * `kernels` = every other kernel of the path (Init/Copy/Mul/Add/Nstream/Dot, reduce u32/f32, heatEquation2D) timed
    the same way, as absolute GB/s and fraction of the HBM roofline;
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


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


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


# ----------------------------------------------------------------------------------------------- reference arm
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def time_reference_triad(n: int, runs: int):
    """Reference TriadKernel on AccCpuOmp2Blocks (oracle/_ref), timed the reference's way (host clock around
    exec + wait, babelStreamMainTest.cpp:279-291). Returns list of seconds per run, threads used."""
    import oracle_lib as ol

    L = ol.ref()
    if L is None:
        return None, 0, "oracle/_ref/libalpaka_ref.so missing"
    secs = (C.c_double * runs)()
    rc = L.ref_babelstream_time(4, 1, n, runs, secs, 0)
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
    secs, threads, err = time_reference_triad(n, runs)
    if secs is None:
        kind = "port"
        secs, threads = time_port_triad(n, runs)
    timed = secs[args.warmup:]
    mean_s = sum(timed) / len(timed)
    value = 24.0 * n * 1e-9 / mean_s
    sample = (f"reference TriadKernel (babelStreamMainTest.cpp:125-141) on AccCpuOmp2Blocks<1,uint32>, double, "
              f"n=2^{n.bit_length() - 1} per step, {threads} OpenMP threads, host clock around exec+wait; "
              f"best step {24.0 * n * 1e-9 / min(timed):.1f} GB/s")
    if kind == "port":
        sample = f"oracle C port (OpenMP) because: {err}; n={n}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BabelStream Triad double, 2^30 elements/array per GPU (BASELINE.json configs[1])",
                   "elements_per_step": n, "bytes_per_element": 24},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
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
    lib = _lib.load()
    peak, peak_kind = measured_peak()

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

    def record(name, ms, nbytes):
        gbs = world * nbytes * 1e-9 / (ms * 1e-3)
        kernels[name] = {"gbs": round(gbs, 1), "ms": round(ms, 4), "frac_of_hbm_peak": round(gbs / world / peak, 4)}

    record("triad_f64", ms_triad, triad_bytes)
    exch = None
    if not args.quick and not args.only_heat:
        Ks = max(5, K // 2)
        record("init_f64", timed(lambda: bs.init(q, a, b, c), Ks, 3), 24.0 * n)
        bs.copy(q, a, b)
        bs.mul(q, a, b)
        record("copy_f64", timed(lambda: bs.copy(q, a, c), Ks, 3), 16.0 * n)
        record("mul_f64", timed(lambda: bs.mul(q, a, b), Ks, 3), 16.0 * n)
        record("add_f64", timed(lambda: bs.add(q, a, b, c), Ks, 3), 24.0 * n)
        record("nstream_f64", timed(lambda: bs.nstream(q, c, a, b, 0.0), Ks, 3), 32.0 * n)
        out = ab.alloc_buf(dev, np.float64, 1, q)
        exch = None
        if dist is None:
            record("dot_f64", timed(lambda: bs.dot_async(q, a, b, out), Ks, 3), 16.0 * n)
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
        out.free()

    for buf in (a, b, c):
        buf.free()
    q.wait()

    if not args.quick:
        if not args.only_heat:
            # ---- example/reduce: 2^32 uint32 (17.2 GB) and 2^30 float (BASELINE.json configs[2])
            nr = (1 << 32) if not args.n else 4 * n
            src = ab.alloc_buf(dev, np.uint32, nr, q)
            res = ab.alloc_buf(dev, np.uint32, 1, q)
            # fill with ones via the f32 init kernel's bit pattern is not exact for u32; memset 0x01010101 instead
            ab.memset(q, src, 1)
            Ks = max(5, K // 2)
            if exch is None:
                record("reduce_u32", timed(lambda: ab.reduce.reduce_sum_async(q, src, res), Ks, 3), 4.0 * nr)
            else:
                record("reduce_u32", timed(lambda: exch.reduce_sum_async(q, src, res), Ks, 3), 4.0 * nr)
                kernels["reduce_u32"]["exchange"] = kernels["dot_f64"]["exchange"]
            got = np.empty(1, dtype=np.uint32)
            ab.memcpy(q, got, res)
            q.wait()
            assert int(got[0]) == (0x01010101 * nr * world) % 2**32, "reduce u32 check failed"
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
        if world == 1:
            dx, dy = 1.0 / (NX + 1), 1.0 / (NY + 1)
            dt = 0.2 * min(dx * dx, dy * dy)
            h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
            # the reference driver's initial condition (analyticalSolution.hpp:58-71), NOT zeros: the power an FP64 kernel
            # draws, and with it the clocks of a long run, depend on the operand bits
            h.upload(ab.heat2d.initial_field(NY, NX, dx, dy))
            record("heat2d_f64", timed(lambda: h.step(1), max(20, K), 5), 16.0 * NY * NX)
            # 2, 3 and 4 time levels per launch (b200_heat2d_step2_f64 / b200_heat2d_stepn_f64): the same 16 B per cell per
            # step of ALGORITHMIC bytes, a half / a third / a quarter of them actually moved, so the fraction of the HBM peak exceeds 1
            for G in (4, 3, 2):  # the default depth first, on a board that is still cool
                record(f"heat2d_f64_{G}_steps_per_launch", timed(lambda: h.step(G, fuse=G), max(20, K), 5), G * 16.0 * NY * NX)
                kG = kernels[f"heat2d_f64_{G}_steps_per_launch"]
                kG["ms_per_step"] = round(kG["ms"] / G, 4)
                # what one launch actually moves: one read + one write of the field (ncu: 4.25 GB at 16384^2), whatever G
                kG["hbm_pass_gbs"] = round(16.0 * NY * NX * 1e-9 / (kG["ms"] * 1e-3), 1)
                kG["hbm_pass_frac_of_peak"] = round(kG["hbm_pass_gbs"] / peak, 4)
            # BASELINE.json configs[3] as specified: 1000 FTCS steps in one go (sustained clocks, not a short burst)
            for G in (() if args.no_sustained else (1, 4)):
                ms_step, clk = timed_run(lambda: h.step(1000, fuse=G), 1000)
                record(f"heat2d_f64_1000_steps_{G}_per_launch", ms_step, 16.0 * NY * NX)
                kernels[f"heat2d_f64_1000_steps_{G}_per_launch"].update(sustained=True, ms_per_step=round(ms_step, 4), clocks=clk)
            h.close()
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
            import math

            sx, sy = ab.heat2d.boundary_tables(tile.ny, tile.nx, runner.dx, runner.dy, tile.j_offset, tile.i_offset)
            exact = math.exp(-math.pi * math.pi * tmax) * (sx[None, :] + sy[:, None])
            err = float(np.max(np.abs(local[1:-1, 1:-1] - exact[1:-1, 1:-1])))
            assert err < 1e-4, f"decomposed heat field deviates from the analytic solution: {err}"
            kernels["heat2d_f64"]["max_abs_error_vs_analytic"] = err
            runner.close()
            # the same field as row slabs advanced TWO time levels per launch and per exchange (ghost rows two deep):
            # algorithmic bytes stay 16 B per cell per step, half of them are moved
            slab = multi.HeatSlab(q, rank, world, NY, NX)
            multi.connect_over_process_group(slab, dist)
            slab.upload(slab.initial_field())
            barrier()
            G = slab.levels
            ms_slab = timed(lambda: slab.step(G), max(50, K), 5)
            assert slab.status() == 0, "heat slab flag wait timed out"
            name_s = f"heat2d_f64_{G}_steps_per_launch"
            record(name_s, ms_slab, G * 16.0 * NY * NX / world)
            kernels[name_s].update(
                scaling="strong", ms_per_step=round(ms_slab / G, 4),
                decomposition=f"{world} row slabs of {NY // world}x{NX}, ghost rows {G} deep, fused P2P halo")
            if not args.no_sustained:
                ms_step, clk = timed_run(lambda: slab.step(1000), 1000)  # C4 as specified: 1000 steps in launches of slab.levels
                assert slab.status() == 0, "heat slab flag wait timed out"
                record(f"heat2d_f64_1000_steps_{G}_per_launch", ms_step, 16.0 * NY * NX / world)
                kernels[f"heat2d_f64_1000_steps_{G}_per_launch"].update(scaling="strong", sustained=True, ms_per_step=round(ms_step, 4), clocks=clk)
            local = slab.download()
            tmax = slab.step_index * slab.dt
            exact = math.exp(-math.pi * math.pi * tmax) * (slab.sx[None, :] + slab.sy[:, None])
            err = float(np.max(np.abs(local[G:-G, 1:-1] - exact[G:-G, 1:-1])))
            assert err < 1e-4, f"slab-decomposed heat field deviates from the analytic solution: {err}"
            kernels[name_s]["max_abs_error_vs_analytic"] = err
            slab.close()
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
                record(f"heat2d_f64_weak_{G}_steps_per_launch", ms_sw, G * 16.0 * NYw * NXw / world)
                kernels[f"heat2d_f64_weak_{G}_steps_per_launch"].update(
                    scaling="weak", ms_per_step=round(ms_sw / G, 4),
                    decomposition=f"{NYw}x{NXw} global, {world} row slabs of {NYw // world}x{NXw}, ghost rows {G} deep")
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
        t_e = (time.perf_counter() - t0) / steps_e
        t_e = max_over_ranks(t_e)
        assert float(hc.array[0]) == 5.0 and float(hc.array[-1]) == 5.0
        e2e = {"value": world * 24.0 * ne * 1e-9 / t_e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "elements_per_step": ne, "steps": steps_e, "ms_per_step": t_e * 1e3,
               "how": "pinned host arrays -> 4-stream chunk pipeline (H2D a,b; Triad; D2H c), wall clock incl. sync"}
        pipe.close()
        del ha, hb, hc

    # ---- CPU baseline beside it (rank 0, N=1 only): reference AccCpuOmp2Blocks Triad at C1's 2^25
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        use_all_host_threads()
        n_cpu, runs = 1 << 25, 21
        secs, threads, err = time_reference_triad(n_cpu, runs)
        kind = "reference"
        if secs is None:
            kind = "port"
            secs, threads = time_port_triad(n_cpu, runs)
        best = min(secs[1:])
        cpu = {"value": 24.0 * n_cpu * 1e-9 / best, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"Triad double n=2^25 (BASELINE.json configs[0]), {runs} runs, min excluding the first "
                         f"(reference method), AccCpuOmp2Blocks" + ("" if kind == "reference" else f" [port: {err}]")}

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
            "config": {"workload": "BabelStream Triad double, 2^30 elements/array per GPU (BASELINE.json configs[1])",
                       "elements_per_gpu": n, "bytes_per_element": 24,
                       "l2": "inputs (25.8 GB per GPU) larger than L2, no flush needed",
                       "parallelism": f"slab x{world}" if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_kind": f"{peak_kind} (burst copy figure; kernel timed alone)",
                         "kernel": "streamKernel<TriadOp<double>>"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "kernels": kernels,
        }
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
    ap.add_argument("--only-heat", action="store_true", help="Triad headline + the heatEquation2D lines only")
    ap.add_argument("--no-sustained", action="store_true", help="skip the 1000-step heat runs (profiling under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
