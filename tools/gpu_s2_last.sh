#!/bin/bash
# round 2, session 2, last call: the final tree — smoke on the prebuilt library + the whole GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
echo "== smoke"; SMOKE_NO_REBUILD=1 timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu_s2_final.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu_s2_final.txt | cut -c1-300
