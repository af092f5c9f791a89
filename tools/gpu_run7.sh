#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
echo "== join"
$B --workload join > gpurun_out/bench_join_dense.json 2>gpurun_out/bench_join_dense.err; python -c "import json; d=json.load(open('gpurun_out/bench_join_dense.json')); print('dense', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()}, d['roofline']['frac'])"
BL_JOIN_DENSE=0 $B --workload join > gpurun_out/bench_join_hash.json 2>gpurun_out/bench_join_hash.err; python -c "import json; d=json.load(open('gpurun_out/bench_join_hash.json')); print('hash', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()}, d['roofline']['frac'])"
echo "== low cardinality"
for k in 1000 2000; do
  $B --keys $k 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('keys=$k', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.05})"
done
echo "== h2d bandwidth"
python - <<'PY'
import time, numpy as np, polars_b200 as plb
plb.init()
a = plb.to_pinned(np.zeros(100_000_000, np.int64))
for i in range(4):
    t=time.perf_counter(); d = plb.to_device(a); dt=time.perf_counter()-t; print("H2D 800MB pinned GB/s", round(0.8/dt,1)); d.free()
b = np.zeros(100_000_000, np.int64)
t=time.perf_counter(); d = plb.to_device(b); dt=time.perf_counter()-t; print("H2D 800MB pageable GB/s", round(0.8/dt,1))
t=time.perf_counter(); v,_ = d.to_numpy(); dt=time.perf_counter()-t; print("D2H 800MB (to pinned + numpy copy) GB/s", round(0.8/dt,1))
PY
echo "== compute-sanitizer memcheck on smoke"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_smoke.log 2>&1; echo "sanitizer exit $?"; grep -E "ERROR SUMMARY|Invalid|smoke ok|Error" gpurun_out/sanitizer_smoke.log | head
