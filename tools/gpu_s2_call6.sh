#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "group_by or config_c1 or large_properties or string" -p no:cacheprovider > gpurun_out/pytest_groupby_s2.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_groupby_s2.log | cut -c1-300
timeout 300 python tools/sweep_bulk.py > gpurun_out/sweep_bulk_v5.jsonl 2> gpurun_out/sweep_bulk_v5.err; echo "sweep rc=$?"; cat gpurun_out/sweep_bulk_v5.jsonl; tail -5 gpurun_out/sweep_bulk_v5.err
