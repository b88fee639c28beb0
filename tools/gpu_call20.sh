#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"; tail -3 $O/bench_n2.err | cut -c1-300
