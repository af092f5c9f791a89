#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sweeps"
for v in "" "BL_K5_LF=30" "BL_K5_LF=15" "BL_K5_BPS=4" "BL_K5_BPS=16" "BL_K5_BPS=32" "BL_K5_LF=30 BL_K5_BPS=16"; do
  echo "-- $v"; env $v $B 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['knobs'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done | tee gpurun_out/sweep_groupby.txt
echo "== join"
$B --workload join > gpurun_out/bench_join_dense.json 2>gpurun_out/bench_join_dense.err; cat gpurun_out/bench_join_dense.json
BL_JOIN_DENSE=0 $B --workload join > gpurun_out/bench_join_hash.json 2>gpurun_out/bench_join_hash.err; cat gpurun_out/bench_join_hash.json
BL_JOIN_DENSE=0 BL_TRACE=1 python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 --workload join 2>&1 | tail -30 > gpurun_out/join_trace.txt
echo "== ncu"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_gb_consume -s 3 -c 1 -o gpurun_out/k5 -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 > gpurun_out/ncu_k5.log 2>&1
timeout 600 $NCU -k regex:"k_join_probe|k_join_emit|k_join_build" -s 9 -c 3 -o gpurun_out/k8 -f env BL_JOIN_DENSE=0 python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 --workload join > gpurun_out/ncu_k8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_groupby.csv python bench.py --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out
