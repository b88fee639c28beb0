#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_heat_halo.py -m gpu -x -q -k "deep" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 240 $TR tools/heat_deep_probe_mp.py 8192 16384 1 2 4 > $O/heat_deep_probe_n2b.log 2>&1; echo "probe rc=$?"; grep "halo_debug" $O/heat_deep_probe_n2b.log
