#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
{
for m in fused8 fused8 fused6 fused4; do build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=1000 --mode=$m | tail -2 | head -1; done
B200_TUNE=heat.walk_minb=3 build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=1000 --mode=fused8 | tail -2 | head -1
python tools/heat_one.py 8 16384 16384 125
nvidia-smi --query-gpu=clocks.sm,clocks_throttle_reasons.active,power.draw --format=csv
} > $O/fused8_probe.log 2>&1
cat $O/fused8_probe.log
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gpu-ref --no-sustained --quick"
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:TriadOp -s 3 -c 1 -f -o $O/triad $BENCH > $O/ncu_triad.log 2>&1; echo "ncu triad rc=$?"
ncu -i $O/triad.ncu-rep --page raw --csv > $O/triad.raw.csv 2>/dev/null
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gpu-ref --no-sustained"
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:MulOp -s 3 -c 1 -f -o $O/mul $BENCH > $O/ncu_mul.log 2>&1; echo "ncu mul rc=$?"
ncu -i $O/mul.ncu-rep --page raw --csv > $O/mul.raw.csv 2>/dev/null
rm -f $O/triad.ncu-rep $O/mul.ncu-rep
