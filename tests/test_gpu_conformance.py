"""The REFERENCE's own unit and integration tests, compiled UNMODIFIED from the reference tree against include/alpaka
(tests/conformance/Makefile), run on the B200 accelerator. This is SURVEY.md section 8f row 1: existing alpaka kernels and
tests build unchanged against the new accelerator and their assertions hold.

Groups (directory under test/ of the reference -> binary): warp (shfl/shfl_up/down/xor, all/any/ballot, activemask,
getSize -- incl. the "half the warp has exited" variants), block/shared (static + dynamic shared memory),
block/sharedSharing, block/sync (+ predicates), idx (mapIdx, mapIdxPitchBytes), workDiv (golden vectors, getValidWorkDiv,
WorkDivMembers), atomic (all ops x types x hierarchies), kernel (lambdas, templates, members, extra params), vec,
intrinsic (popcount, ffs), mem/fence, acc (names, device properties, traits), dev, integ/axpy, integ/sharedMem;
mem/buf, mem/copy (3-D slicing), mem/view (sub-views, ViewConst, plain pointers, device globals), mem/p2p, queue, event,
traits, runtime; exec (uniformElements / uniformGroups / independentGroups / uniformElementsND / oncePerGrid ...:
SURVEY.md section 8f row 3); integ/matMul, integ/mandelbrot, integ/hostOnlyAPI, integ/cudaOnly (CUDA-only mode),
integ/separableCompilation (relocatable device code); the host-only groups unit/meta and unit/core.
The binaries are built where the reference tree exists and travel with the snapshot; a missing binary FAILS."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "conformance")
EXPECTED = ["unit_meta", "unit_core", "integ_axpy", "integ_cudaOnly", "integ_separableCompilation", "integ_hostOnlyAPI", "integ_mandelbrot", "integ_matMul", "integ_sharedMem", "unit_event", "unit_exec",
            "unit_mem_buf", "unit_mem_copy", "unit_mem_p2p", "unit_mem_view", "unit_queue", "unit_runtime", "unit_traits", "unit_acc", "unit_atomic", "unit_block_shared", "unit_block_sharedSharing",
            "unit_block_sync", "unit_dev", "unit_idx", "unit_intrinsic", "unit_kernel", "unit_mem_fence", "unit_vec",
            "unit_warp", "unit_workDiv"]


def manifest():
    path = os.path.join(BIN, "MANIFEST")
    if not os.path.exists(path):
        return EXPECTED
    with open(path) as f:
        names = [ln.strip() for ln in f if ln.strip()]
    return sorted(set(names) | set(EXPECTED))


@pytest.mark.parametrize("name", manifest())
def test_reference_test_group_passes_on_b200(name):
    exe = os.path.join(BIN, name)
    assert os.path.exists(exe), f"{exe} missing: build it with `make -C tests/conformance` where /root/reference exists"
    r = subprocess.run([exe, "--skip-benchmarks"], capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "All tests passed" in r.stdout, tail
