#!/bin/bash
# round-2 GPU call 5b (2 GPUs): ncu of the fused Dot exchange with NVLink counters (one process, two devices), and the
# small-slab regime of the fused heat launch with the exchange attributed (halo_debug knobs)
O=gpurun_out/r02; mkdir -p $O
M=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 python tools/exchange_two_devices.py > $O/exchange_two_devices.log 2>&1; echo "plain run rc=$?"; tail -2 $O/exchange_two_devices.log
timeout 600 ncu --metrics $M --clock-control none --devices 0 -k regex:reduceKernel -s 2 -c 4 --csv --log-file $O/ncu_nvlink_dot_exchange.csv python tools/exchange_two_devices.py > $O/ncu_nvlink_dot_exchange.log 2>&1; echo "ncu dot metrics rc=$?"
timeout 900 ncu --set full --section Nvlink --clock-control none --devices 0 -k regex:reduceKernel -s 2 -c 1 -f -o $O/dot_exchange python tools/exchange_two_devices.py > $O/ncu_full_dot_exchange.log 2>&1; echo "ncu dot full rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR tools/heat_slab_probe.py 4096 16384 4 > $O/heat_slab_probe_n2.log 2>&1; echo "slab probe rc=$?"; grep slab $O/heat_slab_probe_n2.log
