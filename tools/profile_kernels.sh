#!/bin/bash
# tools/profile_kernels.sh -- run on the GPU box (under gpurun): one `ncu --set full` capture per hot-path kernel,
# raw-page CSV extracted next to the report. Usage: tools/profile_kernels.sh <tag> [kernel ...]
# Kernels: triad copy init nstream dot reduce_u32 heat heat2 heat3 heat4 (heatN = N time levels per launch)
set -u
TAG=${1:-r01}; shift || true
KERNELS=${@:-triad copy dot heat}
OUT=gpurun_out/ncu_$TAG
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-sustained"
for k in $KERNELS; do
  case $k in
    triad)      PAT='regex:TriadOp';      SKIP=3; EXTRA="--quick";;
    copy)       PAT='regex:CopyOp';       SKIP=4; EXTRA="";;
    init)       PAT='regex:InitOp';       SKIP=4; EXTRA="";;
    nstream)    PAT='regex:NstreamOp';    SKIP=3; EXTRA="";;
    dot)        PAT='regex:reduceKernel<double'; SKIP=3; EXTRA="";;
    reduce_u32) PAT='regex:reduceKernel<unsigned'; SKIP=3; EXTRA="";;
    heat)       PAT='regex:heatStepKernel'; SKIP=5; EXTRA="";;
    heat2)      PAT='regex:heatStep2Kernel'; SKIP=5; EXTRA="";;
    heat3)      PAT='regex:heatStepNKernel<.int.3'; SKIP=5; EXTRA="";;
    heat4)      PAT='regex:heatStepNKernel<.int.4'; SKIP=5; EXTRA="";;
    *) echo "unknown kernel $k"; continue;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$PAT" -s $SKIP -c 1 -f -o $OUT/$k $BENCH $EXTRA > $OUT/$k.log 2>&1
  ncu -i $OUT/$k.ncu-rep --page raw --csv > $OUT/$k.raw.csv 2>/dev/null
  ncu -i $OUT/$k.ncu-rep --page details --csv > $OUT/$k.details.csv 2>/dev/null
done
ls -la $OUT
