#!/bin/bash
# round-2 GPU call 2 (1 GPU): new tests, full bench line (both arms), fine stream sweep
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_golden_multi.py tests/test_gpu_cpp_layer.py tests/test_gpu_exchange.py tests/test_gpu_heat_halo.py tests/test_gpu_reduce.py -m gpu -x -q > $O/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_new.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 ) > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err; echo "ref rc=$?"
timeout 300 build/tools/stream_lab fine --steps=20 > $O/tune_stream_fine.log 2>&1; echo "lab rc=$?"
