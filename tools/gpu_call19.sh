#!/bin/bash
timeout 150 python -m pytest tests/test_gpu_cpp_layer.py -m gpu -x -q -k "heat2d" 2>&1 | tail -4
