#!/bin/bash
# round 2, call 4 (1 GPU): parity of the bucketized join table + hand-written radix sort, join variants, memcheck of the new kernels, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c4.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c4.txt)
tail -4 gpurun_out/pytest_gpu_c4.txt
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join or tuples or maintain or order or radix" > gpurun_out/memcheck_c4.txt 2>&1; echo rc=$? >> gpurun_out/memcheck_c4.txt)
tail -6 gpurun_out/memcheck_c4.txt
timeout 300 python bench.py --workload join --join-keys sparse --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/join_sparse_w2.json 2> gpurun_out/join_sparse_w2.err
timeout 300 python bench.py --workload join --join-keys sparse --hit-frac 0.5 --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/join_sparse_hit50_w2.json 2> gpurun_out/join_sparse_hit50_w2.err
BL_JOIN_TABLE=compact timeout 300 python bench.py --workload join --join-keys sparse --hit-frac 0.5 --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/join_sparse_hit50_compact.json 2> gpurun_out/join_sparse_hit50_compact.err
timeout 300 python bench.py --workload join --join-keys sparse --dup 4 --no-cpu-baseline --e2e-steps 0 --steps 5 > gpurun_out/join_dup4_w2.json 2> gpurun_out/join_dup4_w2.err
(timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; echo "ref rc=$?")
python - <<'PY'
import json, glob
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return {"ERR": str(e)}
for p in sorted(glob.glob("gpurun_out/join_*w2.json") + glob.glob("gpurun_out/join_*compact.json")):
    j = load(p)
    if "ERR" in j: print(p, j, open(p.replace(".json", ".err")).read()[-600:])
    else: print(p, round(j["ms_per_step"],3), j["verified"][:40], round(j["roofline"]["frac"],4), {k: round(v,3) for k,v in j["kernels_ms_per_step"].items()})
r = load("gpurun_out/bench_ref2.json"); print("REF", json.dumps(r)[:1800])
PY
echo "== radix plan at high cardinality (device-resident, 1e8 rows)"
for keys in 1000000 4000000 10000000 30000000; do
  for mode in 0 2; do
    BL_K5_RADIX=$mode BL_K5_DEBUG=1 timeout 300 python bench.py --workload groupby --keys $keys --no-cpu-baseline --no-verify --e2e-steps 0 --steps 5 2> gpurun_out/radix_${keys}_${mode}.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('keys=$keys radix=$mode', round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernels_ms_per_step'].items() if v > 0.02})"
    grep "k5r\]" gpurun_out/radix_${keys}_${mode}.err | tail -1
  done
done | tee gpurun_out/radix_sweep.txt
