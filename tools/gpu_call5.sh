#!/bin/bash
# round 2, call 5 (1 GPU): full parity suite on the current library, memcheck of the new kernels, join bucket fill A/B, dup emit, K5r sweep, smoke with the source rebuild
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c5.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c5.txt)
tail -15 gpurun_out/pytest_gpu_c5.txt | cut -c1-400
(timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join_vs_oracle or full_join or multi or tuples or radix or deterministic or semi_anti" > gpurun_out/memcheck_c5.txt 2>&1; echo rc=$? >> gpurun_out/memcheck_c5.txt)
tail -6 gpurun_out/memcheck_c5.txt | cut -c1-300
for fill in 1.0 1.4; do
  BL_JOIN_BUCKET_FILL=$fill timeout 300 python bench.py --workload join --join-keys sparse --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/join_sparse_fill$fill.json 2> gpurun_out/join_sparse_fill$fill.err
done
timeout 300 python bench.py --workload join --join-keys sparse --dup 4 --no-cpu-baseline --e2e-steps 0 --steps 5 > gpurun_out/join_dup4_v3.json 2> gpurun_out/join_dup4_v3.err
python - <<'PY'
import json, glob
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return {"ERR": str(e)}
for p in sorted(glob.glob("gpurun_out/join_sparse_fill*.json")) + ["gpurun_out/join_dup4_v3.json"]:
    j = load(p)
    if "ERR" in j: print(p, j, open(p.replace(".json", ".err")).read()[-600:])
    else: print(p, round(j["ms_per_step"],3), str(j["verified"])[:40], round(j["roofline"]["frac"],4), {k: round(v,3) for k,v in j["kernels_ms_per_step"].items()})
PY
echo "== K5r sweep"
for keys in 1000000 4000000 10000000 30000000; do
  for mode in 0 2; do
    BL_K5_RADIX=$mode BL_K5_DEBUG=1 timeout 300 python bench.py --workload groupby --keys $keys --no-cpu-baseline --e2e-steps 0 --steps 5 2> gpurun_out/radix_${keys}_${mode}.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('keys=$keys radix=$mode', round(d['ms_per_step'], 3), str(d['verified'])[:14], {k: round(v, 3) for k, v in d['kernels_ms_per_step'].items() if v > 0.02})"
    grep -E "k5r\]|\[k5\]" gpurun_out/radix_${keys}_${mode}.err | tail -2 | cut -c1-200
  done
done | tee gpurun_out/radix_sweep_v2.txt
echo "== smoke (rebuilds the library from source on this box)"
(time timeout 900 python __graft_entry__.py smoke) > gpurun_out/smoke_c5.txt 2>&1; tail -5 gpurun_out/smoke_c5.txt
