#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log
echo "== default bench"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "import json; d=json.load(open('gpurun_out/bench.json')); print('C2', round(d['ms_per_step'],3), 'value', f\"{d['value']:.3e}\", 'frac', round(d['roofline']['frac'],4), 'e2e ms', round(d['e2e']['ms_per_step'],2), f\"{d['e2e']['value']:.3e}\", 'cpu', f\"{d['cpu_baseline']['value']:.3e}\", d['cpu_baseline']['cores'])"; tail -3 gpurun_out/bench.err
echo "== join (e2e too)"
python bench.py --workload join --no-cpu-baseline > gpurun_out/bench_join.json 2> gpurun_out/bench_join.err; python -c "import json; d=json.load(open('gpurun_out/bench_join.json')); print('C3', round(d['ms_per_step'],3), f\"{d['value']:.3e}\", d['roofline']['kernel'], round(d['roofline']['frac'],3), 'e2e ms', round(d['e2e']['ms_per_step'],2))"; tail -3 gpurun_out/bench_join.err
echo "== q1"
python bench.py --workload q1 --steps 5 > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; python -c "import json; d=json.load(open('gpurun_out/bench_q1.json')); print('Q1', round(d['ms_per_step'],3), f\"{d['value']:.3e}\", {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.02}, 'e2e ms', round(d['e2e']['ms_per_step'],2))"; tail -5 gpurun_out/bench_q1.err
echo "== low card"
for k in 1000 2000; do python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --keys $k 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('keys=$k', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.05})"; done
