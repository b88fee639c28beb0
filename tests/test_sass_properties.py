"""Properties of the COMPILED kernels that the design relies on, checked on the CPU box from the objects and ptxas logs that
__graft_entry__.build() leaves under build/csrc (nvcc cross-compiles for sm_100a without a GPU):

  * no fused multiply-add anywhere in the stream, reduction and stencil kernels -- FMA contraction is pinned OFF on both
    sides of every parity comparison, so a DFMA / FFMA in the SASS would be a parity bug;
  * the stencil tiles really arrive by TMA (UTMALDG) and the N-level kernel exchanges by warp shuffle;
  * the stream kernels move 32-byte vectors (the .256 forms new with sm_100);
  * the default instantiations of the hot kernels do not spill."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "csrc")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def sass(name):
    path = os.path.join(OBJ, name + ".o")
    if not os.path.exists(path) or not os.path.exists(CUOBJDUMP):
        pytest.skip(f"{path} or cuobjdump missing (the GPU box only carries the linked library)")
    return subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, check=True).stdout


@pytest.mark.parametrize("obj", ["b200_stream", "b200_reduce", "b200_heat2d"])
def test_no_fused_multiply_add_in_the_hot_path_objects(obj):
    text = sass(obj)
    assert "DMUL" in text or "FMUL" in text or "DADD" in text  # the object does hold arithmetic
    assert not re.search(r"\b[DFH]FMA\b", text), "contraction must stay off: found an FMA in " + obj


def test_stencil_tiles_arrive_by_tma_and_levels_exchange_by_shuffle():
    text = sass("b200_heat2d")
    assert text.count("UTMALDG") >= 3  # one-level, two-level and N-level kernels
    assert "SHFL.UP" in text and "SHFL.DOWN" in text
    assert "SYNCS" in text or "MBARRIER" in text.upper()  # mbarrier-signalled copies


def test_walker_kernel_loads_by_tma_stores_256_bits_and_exchanges_by_shuffle():
    """heatWalkKernel (the default instantiation at four levels): its stages arrive by TMA (UTMALDG.2D), the quad of a lane
    leaves by one 256-bit store, products travel between lanes by SHFL, there is no CTA-wide barrier (BAR.SYNC) in it, and the
    DP instructions are plain DMUL / DADD (contraction off)."""
    path = os.path.join(OBJ, "b200_heat2d.o")
    if not os.path.exists(path) or not os.path.exists(CUOBJDUMP):
        pytest.skip(f"{path} or cuobjdump missing")
    text = subprocess.run([CUOBJDUMP, "-sass", "-fun", "heatWalkKernel", path], capture_output=True, text=True).stdout
    if "Function" not in text:  # older cuobjdump: no substring match for -fun; cut the functions out of the full dump
        full = sass("b200_heat2d")
        text = "\n".join(blk for blk in full.split("Function : ") if blk.startswith("_ZN") and "heatWalkKernelILi4ELi4ELi4ELb1ELi3ELb0ELb0E" in blk.split("\n")[0])
    else:
        text = "\n".join(blk for blk in text.split("Function : ") if "heatWalkKernelILi4ELi4ELi4ELb1ELi3ELb0ELb0E" in blk.split("\n")[0])
    assert text, "default walker instantiation not found"
    assert "UTMALDG.2D" in text and "SYNCS" in text
    assert re.search(r"STG\.E[.\w]*\.256", text)
    assert "SHFL.UP" in text and "SHFL.DOWN" in text
    assert "BAR.SYNC" not in text, "the walker must not synchronise CTA-wide"
    assert text.count("DADD") > 100 and text.count("DMUL") > 50 and not re.search(r"\bDFMA\b", text)


def test_streams_use_32_byte_vector_accesses():
    text = sass("b200_stream")
    assert re.search(r"LDG\.E[.\w]*\.256", text) and re.search(r"STG\.E[.\w]*\.256", text)


def ptxas_entries(name):
    path = os.path.join(OBJ, name + ".ptxas.log")
    if not os.path.exists(path):
        pytest.skip(f"{path} missing")
    out, cur = {}, None
    for line in open(path):
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            out[cur] = tuple(int(g) for g in m.groups())
    return out


@pytest.mark.parametrize("obj,needles", [
    ("b200_stream", ["TriadOpIdEEdLi32ELi1ELi1E", "CopyOpIdEEdLi32ELi1ELi1E", "NstreamOpIdEEdLi32ELi1ELi1E"]),
    ("b200_reduce", ["reduceKernelIdLb1ELi2E", "reduceKernelIjLb0ELi4E", "reduceKernelIfLb0ELi4E"]),
    ("b200_heat2d", ["heatStepKernelILi1ELi8E", "heatStep2KernelILi1ELi64ELi16E", "heatStepNKernelILi4ELi16ELi2ELb1ELi4E", "heatStepNKernelILi4ELi16ELi2ELb0ELi4E",
                     "heatStepNKernelILi3ELi16ELi2ELb1ELi4E",
                     # the walker kernel's default instantiations: 4 levels (168 registers), 6 and 8 levels (255), square cells
                     "heatWalkKernelILi4ELi4ELi4ELb1ELi3ELb0ELb0E", "heatWalkKernelILi6ELi4ELi4ELb1ELi2ELb0ELb0E"]),
])
def test_default_instantiations_do_not_spill(obj, needles):
    entries = ptxas_entries(obj)
    for needle in needles:
        hits = {k: v for k, v in entries.items() if needle in k}
        assert hits, f"no entry function matching {needle} in {obj}"
        for k, (stack, st, ld) in hits.items():
            assert (stack, st, ld) == (0, 0, 0), f"{k}: {stack} bytes stack, {st}/{ld} bytes spilled"


def test_eight_level_walker_spills_only_outside_its_bare_loop():
    """heatWalkKernel<8, ...> sits at the 255-register ceiling; the few words it spills (integer loop constants) must be
    reloaded in the edge-window and careful paths only: between the first pair of LDS.128 and the fourth 256-bit store after
    it -- the four unrolled rows of the bare chunk -- there is no local-memory access."""
    entries = ptxas_entries("b200_heat2d")
    hits = {k: v for k, v in entries.items() if "heatWalkKernelILi8ELi4ELi4ELb1ELi2ELb0ELb0E" in k}
    assert hits
    for k, (stack, st, ld) in hits.items():
        assert stack <= 96, f"{k}: {stack} bytes of stack"
    full = sass("b200_heat2d")
    blocks = [b for b in full.split("Function : ") if "heatWalkKernelILi8ELi4ELi4ELb1ELi2ELb0ELb0E" in b.split("\n")[0]]
    assert blocks
    lines = blocks[0].splitlines()
    first = next(i for i, l in enumerate(lines) if "LDS.128" in l)
    stores = [i for i, l in enumerate(lines) if i > first and re.search(r"STG\.E[.\w]*\.256", l)]
    assert len(stores) >= 4
    bare = "\n".join(lines[first:stores[3] + 1])
    assert "LDL" not in bare and "STL" not in bare, "the bare loop of the eight-level walker touches local memory"


def test_block_and_grid_hierarchy_atomics_have_device_scope():
    """alpaka::hierarchy::Blocks means "atomic between all blocks of a grid" (reference: atomic/AtomicUniformCudaHip.hpp:80-130
    uses the *_block intrinsics for hierarchy::Threads only): in tests/cpp/atomic_blocks.cpp every Blocks / Grids atomic must
    compile to .GPU scope, and exactly the one hierarchy::Threads atomic to CTA scope (.SM)."""
    path = os.path.join(ROOT, "build", "examples", "test_atomic_blocks")
    if not os.path.exists(path) or not os.path.exists(CUOBJDUMP):
        pytest.skip(f"{path} or cuobjdump missing")
    text = subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = []
        elif cur:
            per_kernel[cur] += re.findall(r"\b(?:ATOMG|REDG|ATOM|RED)\.[\w.]+", line)
    blocks = {k: v for k, v in per_kernel.items() if "CountBlocksKernel" in k}
    threads = {k: v for k, v in per_kernel.items() if "CountThreadsKernel" in k}
    assert blocks and threads  # plain and coarsened trampolines of both functors
    for k, atomics in blocks.items():
        assert len(atomics) >= 6 and all(a.endswith(".GPU") for a in atomics), (k, atomics)  # u32 add x3, f64 add, max, cas
    for k, atomics in threads.items():
        assert atomics and all(a.endswith(".SM") or a.endswith(".CTA") for a in atomics), (k, atomics)
