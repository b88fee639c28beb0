"""The host-only groups of the reference's unit tests (test/unit/meta, test/unit/core: type-list and integral-type
utilities, Interface / ImplementationBase, CallbackThread, ThreadPool, ClipCast, Utility, BoostPredef, OmpSchedule),
compiled UNMODIFIED against include/alpaka by tests/conformance/Makefile. They touch no device, so they run here on the
CPU box too; the same binaries run again with every other group on the B200 (tests/test_gpu_conformance.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "conformance")


@pytest.mark.parametrize("name", ["unit_meta", "unit_core"])
def test_host_only_reference_test_group_passes(name):
    exe = os.path.join(BIN, name)
    assert os.path.exists(exe), f"{exe} missing: build it with `make -C tests/conformance` where /root/reference exists"
    r = subprocess.run([exe, "--skip-benchmarks"], capture_output=True, text=True, timeout=300)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "All tests passed" in r.stdout, tail
