#!/bin/bash
# round 2, call 10 (2 GPUs): multi-GPU parity + the 2-GPU bench line with the one-call group_by step and the join plan chooser
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_gpu_multi_c10.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_multi_c10.txt); tail -3 gpurun_out/pytest_gpu_multi_c10.txt | cut -c1-300
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_g2_v2.json 2> gpurun_out/bench_g2_v2.err; echo "bench g2 rc=$?"
grep -v "^\[nccl\]" gpurun_out/bench_g2_v2.err | tail -5 | cut -c1-400
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return {"ERR": str(e)}
d = load("gpurun_out/bench_g2_v2.json")
if "ERR" in d: print(d)
else:
    print("G2 C2", round(d["ms_per_step"],3), f'{d["value"]:.3e}', "verified:", d["verified"][:40], "e2e ms", round(d["e2e"]["ms_per_step"],2))
    print("   kernels", {k: round(v,3) for k,v in d["kernels_ms_per_step"].items()})
    for s in d.get("secondary", []):
        print("G2 C3", s["config"]["parallelism"][:60], round(s["ms_per_step"],3), f'{s["value"]:.3e}', "verified:", str(s["verified"])[:30], "e2e ms", round(s["e2e"]["ms_per_step"],2))
        print("   kernels", {k: round(v,3) for k,v in s["kernels_ms_per_step"].items()})
        a = s.get("alternative_plan", {}); print("   alt", a.get("parallelism", "")[:50], a.get("ms_per_step"), a.get("unavailable"))
PY
du -sm gpurun_out
