#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (full)"; timeout 1200 python -m pytest tests -m gpu -q --durations=6 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -14 gpurun_out/pytest_gpu.log
echo "== variants"
timeout 900 python tools/bench_variants.py > gpurun_out/variants.jsonl 2> gpurun_out/variants.err; echo "variants exit $?"; cut -c1-400 gpurun_out/variants.jsonl; tail -3 gpurun_out/variants.err
echo "== racecheck (shared-memory hazards) on small tests"
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "kats or smem_plan or edge or filter_kat or (join_vs_oracle and 3000-2000)" > gpurun_out/racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | head -8
echo "== 1e9-row stretch"
timeout 900 python bench.py --no-cpu-baseline --e2e-steps 0 --steps 3 --warmup 3 --rows 1000000000 2>gpurun_out/bench_1e9.err > gpurun_out/bench_1e9.json; python -c "import json; d=json.load(open('gpurun_out/bench_1e9.json')); print('1e9', round(d['ms_per_step'],2), f\"{d['value']:.3e}\", round(d['roofline']['frac'],4), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.1})" || tail -3 gpurun_out/bench_1e9.err
