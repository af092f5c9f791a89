#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
echo "== join fused vs two-pass"
for v in "" "BL_JOIN_FUSED=0" "BL_JOIN_DENSE=0" "BL_JOIN_DENSE=0 BL_JOIN_FUSED=0"; do
  env $v timeout 300 $B --workload join 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.02}, d['roofline']['kernel'], round(d['roofline']['frac'],3))"
done | tee gpurun_out/sweep_join.txt
echo "== C2 + kernels"
timeout 300 $B 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('C2', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"
timeout 600 python tools/bench_kernels.py 2>/dev/null | grep -E "gt i64|lt f64" 
echo "== ncu smem kernel keys=1000 + fused join"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_gb_consume_smem -s 3 -c 1 -o gpurun_out/k5smem -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 --keys 1000 > gpurun_out/ncu_k5smem.log 2>&1
timeout 600 $NCU -k regex:k_join_probe_emit -s 3 -c 1 -o gpurun_out/k8fused -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 --workload join > gpurun_out/ncu_k8fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
