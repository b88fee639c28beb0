#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 200 python tools/heat_seg_probe.py > $O/heat_seg_probe.log 2>&1; echo rc=$?; cat $O/heat_seg_probe.log
