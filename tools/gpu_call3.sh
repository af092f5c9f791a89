#!/bin/bash
# round 2, call 3 (1 GPU): parity suite, the new default bench line, the reference arm, dup-join variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|Thread" >> gpurun_out/host.txt
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c3.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c3.txt)
tail -4 gpurun_out/pytest_gpu_c3.txt
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err; echo "bench rc=$?")
tail -3 gpurun_out/bench_all.err
(timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?")
timeout 300 python bench.py --workload join --join-keys sparse --dup 4 --no-cpu-baseline --e2e-steps 0 --steps 5 > gpurun_out/join_dup4.json 2> gpurun_out/join_dup4.err
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return {"ERR": str(e)}
d = load("gpurun_out/bench_all.json")
if "ERR" in d: print(d)
else:
    print("C2", round(d["ms_per_step"],3), f'{d["value"]:.3e}', "frac", round(d["roofline"]["frac"],4), "verified:", d["verified"], "e2e ms", round(d["e2e"]["ms_per_step"],2))
    print("   kernels", {k: round(v,3) for k,v in d["kernels_ms_per_step"].items()})
    for s in d.get("secondary", []):
        print("C3", s["config"]["workload"][:90]); print("   ", round(s["ms_per_step"],3), f'{s["value"]:.3e}', s["roofline"]["kernel"], "frac", round(s["roofline"]["frac"],4), "verified:", s["verified"], "e2e ms", round(s["e2e"]["ms_per_step"],2))
        print("   kernels", {k: round(v,3) for k,v in s["kernels_ms_per_step"].items()})
    print("cpu_baseline", json.dumps(d.get("cpu_baseline"))[:1500])
r = load("gpurun_out/bench_ref.json"); print("REF", json.dumps(r)[:1200])
j = load("gpurun_out/join_dup4.json")
if "ERR" in j: print(j)
else: print("dup4", round(j["ms_per_step"],3), j["verified"], {k: round(v,3) for k,v in j["kernels_ms_per_step"].items()})
PY
