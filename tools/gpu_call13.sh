#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_heat_halo.py -m gpu -x -q -k "deep" ) > $O/pytest_deep.log 2>&1; echo "pytest rc=$?"; tail -30 $O/pytest_deep.log
