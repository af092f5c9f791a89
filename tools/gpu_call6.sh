#!/bin/bash
# round 2, call 6 (2 GPUs): full parity suite incl. the multi-GPU test, smoke with rebuild, K5r sweep, K6 rate, 2-GPU bench (group_by + join)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L | head -4
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_c6.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c6.txt)
tail -12 gpurun_out/pytest_gpu_c6.txt | cut -c1-400
(time timeout 900 python __graft_entry__.py smoke) > gpurun_out/smoke_c6.txt 2>&1; tail -4 gpurun_out/smoke_c6.txt | cut -c1-300
echo "== K5r sweep"
for keys in 3000000 4000000 10000000; do
  for mode in 0 1; do
    BL_K5_RADIX=$mode BL_K5_DEBUG=1 timeout 300 python bench.py --workload groupby --keys $keys --no-cpu-baseline --e2e-steps 0 --steps 5 2> gpurun_out/radix_${keys}_${mode}.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('keys=$keys radix=$mode', round(d['ms_per_step'], 3), str(d['verified'])[:14], {k: round(v, 3) for k, v in d['kernels_ms_per_step'].items() if v > 0.02})"
    grep -E "k5r\]" gpurun_out/radix_${keys}_${mode}.err | tail -1 | cut -c1-200
  done
done | tee gpurun_out/radix_sweep_v3.txt
echo "== kernels"
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels_c6.jsonl 2> gpurun_out/kernels_c6.err; cut -c1-220 gpurun_out/kernels_c6.jsonl
echo "== 2-GPU bench"
NCCL_DEBUG=INFO timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 rc=$?"
tail -5 gpurun_out/bench_g2.err | cut -c1-400
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return {"ERR": str(e)}
d = load("gpurun_out/bench_g2.json")
if "ERR" in d: print(d)
else:
    print("G2 C2", round(d["ms_per_step"],3), f'{d["value"]:.3e}', "verified:", d["verified"], "e2e ms", round(d["e2e"]["ms_per_step"],2), f'{d["e2e"]["value"]:.3e}')
    print("   kernels", {k: round(v,3) for k,v in d["kernels_ms_per_step"].items()})
    for s in d.get("secondary", []):
        print("G2 C3", s["config"]["workload"][60:150]); print("   ", round(s["ms_per_step"],3), f'{s["value"]:.3e}', "verified:", s["verified"], "e2e ms", round(s["e2e"]["ms_per_step"],2))
        print("   kernels", {k: round(v,3) for k,v in s["kernels_ms_per_step"].items()})
        if "broadcast_variant" in s: print("   broadcast", json.dumps(s["broadcast_variant"])[:600])
PY
