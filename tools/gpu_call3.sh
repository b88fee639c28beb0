#!/bin/bash
# round-2 GPU call 3 (1 GPU): full GPU test suite, bench, C++ drivers (native / renamed / plain generic)
O=gpurun_out/r02; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
{
  echo "## babelstream_b200 native"; build/examples/babelstream_b200 --array-size=1073741824 --number-runs=10 | tail -9
  echo "## babelstream_b200_renamed (generic path, coarsened)"; build/examples/babelstream_b200_renamed --array-size=1073741824 --number-runs=10 | tail -9
  echo "## babelstream_b200_renamed generic.coarsen=0 (plain one-element-per-thread trampoline)"; B200_TUNE=generic.coarsen=0 build/examples/babelstream_b200_renamed --array-size=1073741824 --number-runs=10 | tail -9
  echo "## heat2d_b200 functors native / generic"; build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=100 --mode=functors | tail -2
  ALPAKA_B200_NATIVE=0 build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=100 --mode=functors | tail -2
  echo "## heat2d_b200 fused4"; build/examples/heat2d_b200 --ny=16384 --nx=16384 --steps=1000 --mode=fused4 | tail -2
} > $O/cpp_drivers.log 2>&1
echo "drivers rc=$?"
python - <<'PY' > $O/heat_variants.log 2>&1
import sys, numpy as np
sys.path.insert(0, '.')
import alpaka_b200 as ab
dev = ab.Platform().get_dev_by_idx(0); q = ab.Queue(dev)
NY = NX = 16384
dx = dy = 1.0 / (NX + 1); dt = 0.2 * dx * dx
h = ab.heat2d.Heat2D(q, NY, NX, dx, dy, dt)
h.upload(ab.heat2d.initial_field(NY, NX, dx, dy))
e0, e1 = ab.Event(dev, timing=True), ab.Event(dev, timing=True)
def t(fn, n):
    fn(); q.wait(); ab.enqueue(q, e0)
    for _ in range(n): fn()
    ab.enqueue(q, e1); q.wait(); return e0.elapsed_ms(e1) / n
for name, tune in (("default", {}), ("sq=0", {"heat.stepn_sq": 0}), ("ctas=5", {"heat.stepn_ctas": 5}), ("rpt=32", {"heat.stepn_rpt": 32}), ("nwy=4", {"heat.stepn_nwy": 4})):
    for k, v in tune.items(): ab.runtime.tune_set(k, v)
    burst = t(lambda: h.step(4, fuse=4), 25) / 4
    long_ = t(lambda: h.step(1000, fuse=4), 1) / 1000
    print(f"heat 16384^2 4 levels/launch {name:8s}: burst {burst*1e3:.1f} us/step, 1000 steps {long_*1e3:.1f} us/step", flush=True)
    for k in tune: ab.runtime.tune_set(k, {"heat.stepn_sq": 1, "heat.stepn_ctas": 4, "heat.stepn_rpt": 16, "heat.stepn_nwy": 2}[k])
PY
echo "variants rc=$?"; cat $O/heat_variants.log
