#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
echo "== cardinality sweep (multi-pass)"
for k in 1000000 1500000 3000000 10000000; do timeout 300 $B --keys $k 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('keys=$k', round(d['ms_per_step'],3), {k:(round(v,3)) for k,v in d['kernels_ms_per_step'].items() if v>0.05}, 'launches', d['gpu_launches'])"; done | tee gpurun_out/sweep_card2.txt
BL_K5_MULTIPASS=0 timeout 300 $B --keys 10000000 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('single-pass keys=1e7', round(d['ms_per_step'],3))" | tee -a gpurun_out/sweep_card2.txt
