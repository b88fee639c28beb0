#!/bin/bash
# round-2 GPU call 7 (1 GPU): ncu --set full of the walker kernel, 4 and 8 levels
O=gpurun_out/r02; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:heatWalkKernel -s 2 -c 1 -f -o $O/walk4 python tools/heat_one.py 4 > $O/ncu_walk4.log 2>&1; echo "ncu walk4 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:heatWalkKernel -s 2 -c 1 -f -o $O/walk8 python tools/heat_one.py 8 > $O/ncu_walk8.log 2>&1; echo "ncu walk8 rc=$?"
