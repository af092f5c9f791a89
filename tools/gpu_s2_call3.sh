#!/bin/bash
# round 2, session 2, call 3: why is the bulk-reduce K5 slower than the micro-benchmark predicted?  K5-like ubench variants + ncu of both kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 120 tools/ubench2 > gpurun_out/ubench2_v2.jsonl 2>&1; grep "k5_like\|red_plans" gpurun_out/ubench2_v2.jsonl
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:k_gb_consume -o gpurun_out/r02_k5_bulk -f python tools/ncu_k5.py > gpurun_out/ncu_k5.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_k5.log
ls -la gpurun_out/*.ncu-rep
