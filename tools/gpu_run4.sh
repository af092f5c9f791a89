#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
echo "== low cardinality"
for v in "BL_K5_SMEM=1" "BL_K5_SMEM=0"; do for k in 4 100 1000 3000 10000 100000; do
  env $v $B --keys $k 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v keys=$k', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.05})"
done; done | tee gpurun_out/sweep_lowcard.txt
echo "== default bench (full, with cpu baseline + e2e)"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json
echo "== 1e9-row stretch"
$B --rows 1000000000 --steps 3 2>>gpurun_out/sweep.err > gpurun_out/bench_1e9.json; python -c "import json; d=json.load(open('gpurun_out/bench_1e9.json')); print('1e9', d['ms_per_step'], d['value'], d['roofline']['frac'])"
echo "== ncu"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_gb_consume -s 3 -c 1 -o gpurun_out/k5 -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 > gpurun_out/ncu_k5.log 2>&1
timeout 600 $NCU -k regex:"k_join_dense_probe|k_join_emit" -s 6 -c 2 -o gpurun_out/k8 -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 --workload join > gpurun_out/ncu_k8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_groupby.csv python bench.py --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ls gpurun_out
