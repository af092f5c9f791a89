#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
echo "== cardinality sweep"
for k in 4 100 1000 1500000 3000000; do timeout 300 $B --keys $k 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('keys=$k', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.05})"; done | tee gpurun_out/sweep_card.txt
echo "== q1"; timeout 300 python bench.py --workload q1 --steps 5 --e2e-steps 1 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('Q1', round(d['ms_per_step'],3), f\"{d['value']:.3e}\", {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.02})"
echo "== ncu K5 SoA"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gb_consume -s 3 -c 1 -o gpurun_out/k5soa -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 > gpurun_out/ncu_k5.log 2>&1; tail -1 gpurun_out/ncu_k5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_groupby.csv python bench.py --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
