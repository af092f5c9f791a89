#!/bin/bash
# round 2, session 2, call 4: ncu of the bulk-reduce K5 and the 3-RED K5; pages exported on the box (the report itself is > 64 MiB: the module's cubin is embedded per result)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --kernel-name regex:k_gb_consume -o /tmp/r02_k5_bulk -f python tools/ncu_k5.py > gpurun_out/ncu_k5.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_k5.log
ncu -i /tmp/r02_k5_bulk.ncu-rep --page raw --csv > gpurun_out/r02_k5_bulk_raw.csv 2>/dev/null
ncu -i /tmp/r02_k5_bulk.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r02_k5_bulk_sass.csv.gz
ncu -i /tmp/r02_k5_bulk.ncu-rep --page details 2>/dev/null | gzip > gpurun_out/r02_k5_bulk_details.txt.gz
ls -la gpurun_out /tmp/r02_k5_bulk.ncu-rep
