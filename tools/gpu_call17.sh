#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cpp_layer.py tests/test_gpu_heat_halo.py -m gpu -x -q -k "tiles or deep" 2>&1 | tail -8
