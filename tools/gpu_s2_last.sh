#!/bin/bash
# round 2, session 2, last call: the final tree — the whole GPU suite, then smoke() with its forced rebuild from source on this box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu_s2_final.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu_s2_final.txt | cut -c1-300
echo "== smoke (rebuilds the library from source)"; /usr/bin/time -v timeout 330 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_s2.txt 2>&1; echo "smoke rc=$?"; grep -E "smoke ok|Elapsed|Error|error" gpurun_out/r02_smoke_s2.txt | head -5
