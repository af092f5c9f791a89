#!/bin/bash
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "chunked_sliced or error_paths or null_scalar or kats or filter" 2>&1 | tail -6
echo "== ubench SoA"; timeout 300 tools/ubench 2>&1 | grep -E "SoA|red_random" | grep -E "SoA|\"log2_entries\": 2[012]"
echo "== ncu K5 (32-byte entries)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gb_consume -s 3 -c 1 -o gpurun_out/k5 -f python bench.py --no-cpu-baseline --e2e-steps 0 --steps 1 --warmup 3 > gpurun_out/ncu_k5.log 2>&1; tail -2 gpurun_out/ncu_k5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_groupby.csv python bench.py --no-cpu-baseline --e2e-steps 0 --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
