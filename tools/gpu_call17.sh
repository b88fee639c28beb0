#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_heat2d.py tests/test_gpu_heat_halo.py tests/test_gpu_cpp_layer.py tests/test_golden_multi.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/heat_depth_probe.py 2>&1 | tail -5
