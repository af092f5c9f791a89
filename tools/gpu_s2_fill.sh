#!/bin/bash
# round 2, session 2: hashed join probe vs bucket fill (is the probe bound by second-bucket DRAM fetches?)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for f in 0.5 0.7; do
  BL_JOIN_BUCKET_FILL=$f timeout 150 python bench.py --workload join --join-keys sparse --no-cpu-baseline --e2e-steps 1 --steps 6 $([ "$f" = "0.7" ] && echo --no-verify) > gpurun_out/r02_join_sparse_fill$f.json 2> gpurun_out/join_fill$f.err
  python - "$f" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r02_join_sparse_fill{sys.argv[1]}.json"))
print("fill", sys.argv[1], round(d["ms_per_step"], 3), d.get("verified"), d["kernels_ms_per_step"])
PY
done
