#!/bin/bash
# One GPU-box pass: micro-benchmarks, smoke, GPU parity tests, bench line.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
python -c "import polars; print('polars', polars.__version__)" >> gpurun_out/gpu.txt 2>&1
if [ "$1" != "noubench" ]; then timeout 300 tools/ubench > gpurun_out/ubench.jsonl 2>&1; fi
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench join"; timeout 600 python bench.py --workload join --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_join.json 2> gpurun_out/bench_join.err; echo "bench join exit $?"; cat gpurun_out/bench_join.json; tail -5 gpurun_out/bench_join.err
