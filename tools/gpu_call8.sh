#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 300 python tools/heat_depth_probe.py > $O/heat_depth_probe.log 2>&1; echo "rc=$?"; cat $O/heat_depth_probe.log
