#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 300 python tools/heat_deep_probe.py > $O/heat_deep_probe.log 2>&1; echo "probe rc=$?"; cat $O/heat_deep_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_deep.csv python tools/heat_deep_probe.py 2 > $O/launches_deep.log 2>&1; echo "ncu rc=$?"
cut -d, -f5,15 $O/launches_deep.csv | sed 's/(CUtensorMap.*)"/"/; s/void <unnamed>:://' | tail -60
