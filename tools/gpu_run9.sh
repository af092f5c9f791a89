#!/bin/bash
mkdir -p gpurun_out
echo "== plugin tests"; timeout 600 python -m pytest tests/test_gpu_plugin_abi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15
echo "== kernel sweep"; timeout 900 python tools/bench_kernels.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; cat gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
echo "== ncu all kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_arith|k_compare|k_compact|k_gather|k_part_scatter|k_part_count|k_mask_tile" -c 30 -o gpurun_out/kernels -f python tools/bench_kernels.py --once > gpurun_out/ncu_kernels.log 2>&1; tail -3 gpurun_out/ncu_kernels.log
ls -la gpurun_out/*.ncu-rep
