#!/bin/bash
# round 2, session 2, last call: smoke() with its forced rebuild from source on this box (the suite ran in the call before: 262 passed)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
echo "== smoke (rebuilds the library from source)"; SECONDS=0; timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_s2.txt 2>&1; echo "smoke rc=$? seconds=$SECONDS" | tee -a gpurun_out/r02_smoke_s2.txt; tail -3 gpurun_out/r02_smoke_s2.txt | cut -c1-300
