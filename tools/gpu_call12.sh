#!/bin/bash
# round-2 GPU call 12 (N GPUs, N = $1): multi-process parity worker and bench at N with the parity block
N=${1:-8}
O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 300 $TR tests/mp_worker.py ) > $O/mp_worker_n$N.log 2>&1; echo "mp_worker rc=$?"; grep -E "MP_WORKER|Error|error" $O/mp_worker_n$N.log | cut -c1-700 | tail -4
( time timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"; tail -3 $O/bench_n$N.err
nvidia-smi topo -m > $O/topo_n$N.log 2>&1
