#!/bin/bash
# round-2 GPU call 6 (1 GPU): parity of the walker kernel (stand-alone fields and slabs), then its speed against the tile kernel
O=gpurun_out/r02; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_heat2d.py tests/test_gpu_heat_halo.py -m gpu -x -q ) > $O/pytest_walk.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_walk.log
timeout 600 python tools/heat_walk_probe.py > $O/heat_walk_probe.log 2>&1; echo "probe rc=$?"; cat $O/heat_walk_probe.log
LONG_STEPS=240 timeout 300 python tools/heat_walk_probe.py 2048 16384 > $O/heat_walk_probe_2048.log 2>&1; echo "probe2 rc=$?"; cat $O/heat_walk_probe_2048.log
