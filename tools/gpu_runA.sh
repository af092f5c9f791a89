#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== pytest group_by subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "group_by or config_c1 or join_vs_oracle" 2>&1 | tail -3
echo "== low cardinality"
for k in 4 100 1000; do timeout 300 $B --keys $k 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('keys=$k', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.05})"; done | tee gpurun_out/sweep_lowcard.txt
echo "== q1"; timeout 300 python bench.py --workload q1 --steps 5 --e2e-steps 1 > gpurun_out/bench_q1.json 2>>gpurun_out/sweep.err; python -c "import json; d=json.load(open('gpurun_out/bench_q1.json')); print('Q1', round(d['ms_per_step'],3), f\"{d['value']:.3e}\", {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.02})"
echo "== default bench + reference arm + join"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "import json; d=json.load(open('gpurun_out/bench.json')); print('C2', round(d['ms_per_step'],3), f\"{d['value']:.3e}\", 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'e2e', round(d['e2e']['ms_per_step'],2), f\"{d['e2e']['value']:.3e}\", 'cpu', f\"{d['cpu_baseline']['value']:.3e}\", d['cpu_baseline']['cores'], d['clocks'])"; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>>gpurun_out/bench.err; python -c "import json; d=json.load(open('gpurun_out/bench_reference.json')); print('REF', f\"{d['value']:.3e}\", d['cpu_baseline']['cores'])"
python bench.py --workload join --no-cpu-baseline > gpurun_out/bench_join.json 2>>gpurun_out/bench.err; python -c "import json; d=json.load(open('gpurun_out/bench_join.json')); print('C3', round(d['ms_per_step'],3), f\"{d['value']:.3e}\", d['roofline']['kernel'], round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],2))"
