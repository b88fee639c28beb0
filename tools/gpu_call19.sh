#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_heat2d.py -m gpu -x -q -k "not_32_byte" 2>&1 | tail -12
