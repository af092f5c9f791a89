#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest (SoA default)"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
echo "== pytest group_by subset (AoS)"; BL_K5_SOA=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "group_by" 2>&1 | tail -3
echo "== C2 SoA vs AoS"
for v in "BL_K5_SOA=1" "BL_K5_SOA=0" "BL_K5_SOA=1 BL_K5_PAIRS=2" "BL_K5_SOA=1 BL_K5_LF=30" "BL_K5_SOA=1 BL_K5_BPS=16"; do
  env $v timeout 300 $B 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['knobs'], round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4))"
done | tee gpurun_out/sweep_groupby.txt
for k in 1000 10000 100000; do timeout 300 $B --keys $k 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('keys=$k', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.05})"; done
