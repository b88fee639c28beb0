"""The index walks behind alpaka::uniformElements / uniformGroups / uniformGroupElements / independentGroups /
independentGroupElements (include/alpaka/b200/Exec.hpp) against a Python restatement of the reference's iterators
(include/alpaka/exec/UniformElements.hpp:127-190, 703-760, 972-1030; IndependentElements.hpp:70-130, 265-330), on the
host: the ranges are plain value types, so a g++ build without CUDA exercises the same code the kernels run. The device
side (the accessor-derived constructors, uniformElementsND) is covered by the reference's own test/unit/exec on the GPU."""
import itertools
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "tmp", "exec_ranges_host")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "exec_ranges_host.cpp")
    cmd = ["g++", "-std=c++20", "-O1", f"-I{ROOT}/include", "-I/usr/local/cuda/include", src, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return EXE


def ref_elements(first, elements, stride, extent):
    """UniformElementsAlong::const_iterator (UniformElements.hpp:131-165): `elements` consecutive indices, then on by the
    grid stride, clamped to the extent; begin = min(first, extent), end = extent."""
    out, index, index_elem, hop = [], min(first, extent), 0, stride - elements
    while index != extent:
        out.append(index)
        index_elem += 1
        index += 1
        if index_elem >= elements:
            index_elem = 0
            index += hop
        if index >= extent:
            index = extent
    return out


def ref_groups(first, stride, extent):
    """UniformGroupsAlong::const_iterator (UniformElements.hpp:707-732)."""
    out, g = [], min(first, extent)
    while g != extent:
        out.append(g)
        g += stride
        if g >= extent:
            g = extent
    return out


def ref_group_elements(origin, lo, hi):
    """UniformGroupElementsAlong::const_iterator (UniformElements.hpp:976-1001): local indices [lo, hi), global = origin + local."""
    return [(origin + i, i) for i in range(lo, hi)]


def ask(exe, idx_type, lines):
    r = subprocess.run([exe, idx_type], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert r.returncode == 0
    return [ln.split() for ln in r.stdout.split("\n")[: len(lines)]]


@pytest.mark.parametrize("idx_type", ["u32", "i32", "u64"])
def test_run_hop_and_hop_ranges_match_the_reference_iterators(exe, idx_type):
    cases, want = [], []
    # grid of (thread start, elements per thread, grid stride in elements, extent): elements <= stride always
    for elements, threads, extent in itertools.product((1, 2, 3, 8), (1, 2, 5, 32), (0, 1, 7, 64, 100, 257)):
        stride = elements * threads
        for t in sorted({0, 1, threads - 1, threads // 2}):
            for first_off in (0, 3):
                start = t * elements + first_off
                cases.append(f"R {start} {elements} {stride} {extent}")
                want.append([str(v) for v in ref_elements(start, elements, stride, extent)])
    for blocks, extent in itertools.product((1, 2, 7, 148), (0, 1, 6, 7, 300)):
        for b in sorted({0, blocks - 1, blocks // 2}):
            cases.append(f"H {b} {blocks} {extent}")
            want.append([str(v) for v in ref_groups(b, blocks, extent)])
    got = ask(exe, idx_type, cases)
    for c, g, w in zip(cases, got, want):
        assert g == w, c


def test_group_ranges_yield_global_and_local_indices(exe):
    cases, want = [], []
    for block_elems, thread_elems, extent in itertools.product((1, 4, 32), (1, 2, 4), (5, 64, 100)):
        for group in (0, 1, extent // block_elems):
            origin = group * block_elems
            for thread in (0, 1, block_elems // thread_elems - 1 if block_elems >= thread_elems else 0):
                # UniformGroupElementsAlong(acc, group, extent) (UniformElements.hpp:948-957): both ends clipped to extent - origin
                lo = min(max(extent - origin, 0), thread * thread_elems)
                hi = min(max(extent - origin, 0), thread * thread_elems + thread_elems)
                cases.append(f"G {origin} {lo} {hi}")
                want.append([f"{g}:{l}" for g, l in ref_group_elements(origin, lo, hi)])
    got = ask(exe, "u32", cases)
    for c, g, w in zip(cases, got, want):
        assert g == w, c


def test_every_element_is_visited_exactly_once_by_a_grid(exe):
    """The property user kernels rely on: the union over all threads of a launch covers [0, extent) exactly once."""
    elements, threads, extent = 3, 10, 1003
    stride = elements * threads
    lines = [f"R {t * elements} {elements} {stride} {extent}" for t in range(threads)]
    seen = sorted(int(v) for row in ask(exe, "u32", lines) for v in row)
    assert seen == list(range(extent))
