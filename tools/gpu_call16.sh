#!/bin/bash
# round-2 GPU call 16 (1 GPU): final state -- whole GPU suite, smoke(), bench (ours + reference arm)
O=gpurun_out/r02; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 ) > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err; echo "ref rc=$?"
du -sh gpurun_out
