#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (full)"; timeout 1200 python -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== variants"
BL_K5_DEBUG=1 timeout 900 python tools/bench_variants.py > gpurun_out/variants.jsonl 2> gpurun_out/variants.err; echo "variants exit $?"; cut -c1-330 gpurun_out/variants.jsonl; grep "\[k5\]" gpurun_out/variants.err | sort | uniq -c | cut -c1-250; grep -v "\[k5\]" gpurun_out/variants.err | tail -3
