"""The C++ layer's launch scheduler (alpaka::b200::heatNextDepth, include/alpaka/b200/Heat2D.hpp -- what
Heat2DStepper::steps and Heat2DSlabs::steps walk through) against the Python mirror's decomp.launch_schedule, on the host:
both must cover the steps with supported depths only, never deeper than asked, never leave a single step to a slab, and
need the same number of launches (tests/test_gpu_cpp_layer.py compares the drivers' launch counts on the GPU)."""
import os
import subprocess

import pytest

from alpaka_b200 import decomp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "tmp", "heat_schedule_host")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "heat_schedule_host.cpp")
    cmd = ["g++", "-std=c++20", "-O1", f"-I{ROOT}/include", "-I/usr/local/cuda/include", src, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return EXE


def ask(exe, cases):
    r = subprocess.run([exe], input="".join(f"{n} {d} {m}\n" for n, d, m in cases), capture_output=True, text=True)
    assert r.returncode == 0
    return [line.split() for line in r.stdout.splitlines()]


def test_cpp_schedule_matches_the_python_mirror(exe):
    cases = [(n, d, m) for m in (1, 2) for d in (1, 2, 3, 4, 5, 6, 7, 8) if d >= m for n in list(range(0, 70)) + [1000, 1001, 4000]]
    got = ask(exe, cases)
    assert len(got) == len(cases)
    for (n, d, m), sched in zip(cases, got):
        try:
            want = decomp.launch_schedule(n, d, m)
        except ValueError:
            assert sched and sched[-1] == "X", (n, d, m, sched)
            continue
        assert "X" not in sched, (n, d, m, sched)
        ks = [int(k) for k in sched]
        assert sum(ks) == n and all(k in decomp.SUPPORTED_DEPTHS and m <= k <= d for k in ks), (n, d, m, ks)
        assert ks == sorted(ks, reverse=True)
        assert len(ks) == len(want), (n, d, m, ks, want)
