#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python tools/sweep_bulk.py > gpurun_out/sweep_bulk_v3.jsonl 2> gpurun_out/sweep_bulk_v3.err; echo "sweep rc=$?"; cat gpurun_out/sweep_bulk_v3.jsonl; tail -5 gpurun_out/sweep_bulk_v3.err
