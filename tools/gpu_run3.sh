#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sweeps"
for v in "BL_K5_PAIRS=2 BL_K5_HINT=1" "BL_K5_PAIRS=2 BL_K5_HINT=0" "BL_K5_PAIRS=1 BL_K5_HINT=1" "BL_K5_PAIRS=1 BL_K5_HINT=0" "BL_K5_PAIRS=2 BL_K5_HINT=1 BL_K5_BPS=3" "BL_K5_PAIRS=2 BL_K5_HINT=1 BL_K5_BPS=16" "BL_K5_PAIRS=2 BL_K5_HINT=1 BL_K5_LF=25"; do
  env $v $B 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['knobs'], round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4))"
done | tee gpurun_out/sweep_groupby.txt
echo "== join"
$B --workload join > gpurun_out/bench_join_dense.json 2>gpurun_out/bench_join_dense.err; python -c "import json; d=json.load(open('gpurun_out/bench_join_dense.json')); print('dense', d['ms_per_step'], d['kernels_ms_per_step'])"
BL_JOIN_DENSE=0 $B --workload join > gpurun_out/bench_join_hash.json 2>gpurun_out/bench_join_hash.err; python -c "import json; d=json.load(open('gpurun_out/bench_join_hash.json')); print('hash', d['ms_per_step'], d['kernels_ms_per_step'])"
ls gpurun_out | head -30
