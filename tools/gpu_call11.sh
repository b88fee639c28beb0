#!/bin/bash
# round-2 GPU call 11 (2 GPUs): multi-process parity worker, bench at N=2 with the parity block, ncu of the fused Dot
# exchange with NVLink counters (one process, two devices, device 1 enqueued first), walker-kernel slabs: ncu with peer
# stores + the small-slab probe
O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 $TR tests/mp_worker.py ) > $O/mp_worker_n2.log 2>&1; echo "mp_worker rc=$?"; grep -E "MP_WORKER|Error|error" $O/mp_worker_n2.log | cut -c1-600 | tail -4
( time timeout 1200 $TR bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"; tail -3 $O/bench_n2.err
M=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 120 python tools/exchange_two_devices.py > $O/exchange_two_devices.log 2>&1; echo "plain exchange rc=$?"; tail -1 $O/exchange_two_devices.log
timeout 180 ncu --metrics $M --clock-control none --devices 0 -k regex:reduceKernel -s 2 -c 4 --csv --log-file $O/ncu_nvlink_dot_exchange.csv python tools/exchange_two_devices.py > $O/ncu_nvlink_dot_exchange.log 2>&1; echo "ncu dot metrics rc=$?"; tail -2 $O/ncu_nvlink_dot_exchange.log
timeout 240 ncu --set full --section Nvlink --clock-control none --devices 0 -k regex:reduceKernel -s 2 -c 1 -f -o $O/dot_exchange python tools/exchange_two_devices.py > $O/ncu_full_dot_exchange.log 2>&1; echo "ncu dot full rc=$?"
ncu -i $O/dot_exchange.ncu-rep --page raw --csv > $O/dot_exchange.raw.csv 2>/dev/null; rm -f $O/dot_exchange.ncu-rep
HEAT="build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=48 --mode=slabs --slabs=2 --levels=4"
timeout 240 ncu --metrics $M --clock-control none --devices 0 -k regex:heatWalkKernel -s 4 -c 6 --csv --log-file $O/ncu_nvlink_heat_slabs.csv $HEAT > $O/ncu_nvlink_heat_slabs.log 2>&1; echo "ncu heat metrics rc=$?"
timeout 300 ncu --set full --section Nvlink --clock-control none --devices 0 -k regex:heatWalkKernel -s 4 -c 1 -f -o $O/walk4_slab_peer $HEAT > $O/ncu_full_heat_slabs.log 2>&1; echo "ncu heat full rc=$?"
ncu -i $O/walk4_slab_peer.ncu-rep --page raw --csv > $O/walk4_slab_peer.raw.csv 2>/dev/null; rm -f $O/walk4_slab_peer.ncu-rep
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR2 tools/heat_slab_probe.py 4096 16384 4 > $O/heat_slab_probe_n2.log 2>&1; echo "slab probe rc=$?"; grep slab $O/heat_slab_probe_n2.log
nvidia-smi topo -m > $O/topo_n2.log 2>&1
du -sh gpurun_out
