#!/bin/bash
# round 2, call 2: pass-2 prototype variants + GPU parity of the rewritten join + join table A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(timeout 240 tools/proto_radix2 100000000 1000000 > gpurun_out/proto_radix2_v2.log 2>&1; echo rc=$? >> gpurun_out/proto_radix2_v2.log)
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c2.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c2.txt)
for tb in compact wide; do
  BL_JOIN_TABLE=$tb BL_JOIN_DENSE=0 timeout 300 python bench.py --workload join --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/join_${tb}_hashed.json 2> gpurun_out/join_${tb}_hashed.err
  BL_JOIN_TABLE=$tb timeout 300 python bench.py --workload join --dup 4 --no-cpu-baseline --e2e-steps 0 --steps 5 > gpurun_out/join_${tb}_dup4.json 2> gpurun_out/join_${tb}_dup4.err
done
timeout 300 python bench.py --workload join --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/join_dense.json 2> gpurun_out/join_dense.err
tail -5 gpurun_out/pytest_gpu_c2.txt
cat gpurun_out/proto_radix2_v2.log
for f in gpurun_out/join_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], {k: round(v,3) for k,v in d["kernels_ms_per_step"].items()})
except Exception as e:
    print("ERR", e)
PY
done
