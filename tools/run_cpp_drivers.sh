#!/bin/bash
# tools/run_cpp_drivers.sh -- run the C++ alpaka-layer drivers on the GPU box, native kernels on and off.
cd /root/repo/build/examples || exit 1
set -x
./ref_heatEquation2D; echo rc=$?
ALPAKA_B200_NATIVE=0 ./ref_heatEquation2D | tail -1
./ref_babelstream --array-size=33554432 --number-runs=50 2>&1 | grep -A7 -E "Precision|passed|failed"
./ref_babelstream --array-size=1073741824 --number-runs=10 2>&1 | grep -A7 -E "Precision|passed|failed"
ALPAKA_B200_NATIVE=0 ./ref_babelstream --array-size=1073741824 --number-runs=10 2>&1 | grep -A7 -E "Precision:double|passed|failed"
./babelstream_b200 --array-size=1073741824 --number-runs=10 | tail -9
ALPAKA_B200_NATIVE=0 ./babelstream_b200 --array-size=1073741824 --number-runs=10 | tail -2
./reduce_b200 --n=268435456; echo rc=$?
./reduce_b200 --n=4294967296 --runs=5; echo rc=$?
./reduce_b200 --n=1073741824 --dtype=f32 --runs=5; echo rc=$?
./heat2d_b200 --ny=64 --nx=64 --steps=4000 --dt=0.000025 --mode=functors
./heat2d_b200 --ny=64 --nx=64 --steps=4000 --dt=0.000025 --mode=fused
./heat2d_b200 --ny=16384 --nx=16384 --steps=200 --mode=functors
./heat2d_b200 --ny=16384 --nx=16384 --steps=200 --mode=fused
ALPAKA_B200_NATIVE=0 ./heat2d_b200 --ny=16384 --nx=16384 --steps=100 --mode=functors
./heat2d_b200 --ny=16384 --nx=16384 --steps=300 --mode=fused2
./heat2d_b200 --ny=16384 --nx=16384 --steps=300 --mode=fused3
./heat2d_b200 --ny=16384 --nx=16384 --steps=300 --mode=slabs --slabs=1 --levels=3
for e in heatEquation vectorAdd convolution1D convolution2D parallelLoopPatterns; do ./ref_ex_$e | tail -2; done
