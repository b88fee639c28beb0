#!/bin/bash
# round-2 GPU call 1: stream variants sweep, reference CUDA back-end on the same box, host topology, heat ncu capture
mkdir -p gpurun_out/r02
O=gpurun_out/r02
( nvidia-smi topo -m; echo; lscpu | head -40; echo; numactl -H 2>/dev/null || cat /sys/devices/system/node/node*/meminfo | grep MemTotal; echo; free -g; nvidia-smi -q | grep -iE "Link Width|Link Gen|PCIe Generation|Max|Current" | head -20 ) > $O/host_topology.log 2>&1
timeout 600 build/tools/stream_lab --steps=20 > $O/tune_stream_lab.log 2>&1
echo "stream_lab rc=$?"
timeout 600 oracle/_ref/ref_gpu_babelstream --array-size=1073741824 --number-runs=20 > $O/ref_gpu_babelstream.log 2>&1
echo "ref_gpu_babelstream rc=$?"
timeout 300 oracle/_ref/ref_gpu heat 16384 16384 200 > $O/ref_gpu_heat.log 2>&1
echo "ref_gpu heat rc=$?"
timeout 300 oracle/_ref/ref_gpu reduce_u32 4294967296 10 > $O/ref_gpu_reduce.log 2>&1
timeout 300 oracle/_ref/ref_gpu reduce_f32 1073741824 10 >> $O/ref_gpu_reduce.log 2>&1
echo "ref_gpu reduce rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:heatStepNKernel<.int.4" -s 5 -c 1 -f -o $O/heat4 python bench.py --only-heat --no-sustained --no-e2e --no-cpu --steps 3 --warmup 3 > $O/ncu_heat4.log 2>&1
echo "ncu heat rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
tail -5 $O/tune_stream_lab.log
