#!/bin/bash
# round-2 GPU call 9 (1 GPU): the whole GPU suite, bench (ours + reference arm), C++ drivers, launch list of the bench command,
# ncu --set full of the walker (8 and 4 levels), Triad, Mul, Dot at HEAD
O=gpurun_out/r02; mkdir -p $O
if [ -z "$SKIP_PYTEST" ]; then ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log; fi
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
{
  echo "## babelstream_b200 native"; build/examples/babelstream_b200 --array-size=1073741824 --number-runs=10 | tail -9
  echo "## babelstream_b200_renamed (generic path, coarsened)"; build/examples/babelstream_b200_renamed --array-size=1073741824 --number-runs=10 | tail -9
  echo "## heat2d_b200 functors native / generic"; build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=100 --mode=functors | tail -2
  echo "## heat2d_b200 fused4 / fused8 (walker kernel)"; build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=1000 --mode=fused4 | tail -2
  build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=1000 --mode=fused8 | tail -2
} > $O/cpp_drivers.log 2>&1
echo "drivers rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-gpu-ref --no-e2e --no-sustained > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
for spec in "walk8:8" "walk4:4"; do
  name=${spec%%:*}; lv=${spec##*:}
  timeout 300 ncu --set full --clock-control none -k regex:heatWalkKernel -s 2 -c 1 -f -o $O/$name python tools/heat_one.py $lv > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
done
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gpu-ref --no-sustained --quick"
timeout 300 ncu --set full --clock-control none -k regex:TriadOp -s 3 -c 1 -f -o $O/triad $BENCH > $O/ncu_triad.log 2>&1; echo "ncu triad rc=$?"
ncu -i $O/triad.ncu-rep --page raw --csv > $O/triad.raw.csv 2>/dev/null
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gpu-ref --no-sustained"
timeout 300 ncu --set full --clock-control none -k regex:MulOp -s 3 -c 1 -f -o $O/mul $BENCH > $O/ncu_mul.log 2>&1; echo "ncu mul rc=$?"
ncu -i $O/mul.ncu-rep --page raw --csv > $O/mul.raw.csv 2>/dev/null
rm -f $O/walk4.ncu-rep $O/triad.ncu-rep $O/mul.ncu-rep   # keep one report (8 levels); the raw pages of the others are in the CSVs
du -sh gpurun_out
ls -la $O | tail -30
