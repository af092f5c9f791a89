#!/bin/bash
# round 2, call 11 (8 GPUs): sanity of every N = 8 code path at reduced size (the driver runs the full-size scaling bench itself)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --rows 30000000 --keys 1000000 --build-rows 3000000 --e2e-steps 1 > gpurun_out/bench_g8_small.json 2> gpurun_out/bench_g8_small.err; echo "bench g8 rc=$?"
grep -v "^\[nccl\]" gpurun_out/bench_g8_small.err | tail -5 | cut -c1-400; grep -c "^\[nccl\]" gpurun_out/bench_g8_small.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_g8_small.json").read().strip().splitlines()[-1])
    print("G8 C2", round(d["ms_per_step"],3), f'{d["value"]:.3e}', d["verified"][:40], {k: round(v,3) for k,v in d["kernels_ms_per_step"].items()})
    for s in d.get("secondary", []):
        print("G8 C3", s["config"]["parallelism"][:50], round(s["ms_per_step"],3), f'{s["value"]:.3e}', str(s["verified"])[:30], (s.get("alternative_plan") or {}).get("ms_per_step"))
except Exception as e: print("ERR", e)
PY
rm -f gpurun_out/nccl_debug.*.log; du -sm gpurun_out
