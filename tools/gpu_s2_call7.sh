#!/bin/bash
# round 2, session 2, call 7: (a) lean K5 kernel in multi-pass mode (nullable C2 variant), (b) cudaLimitMaxL2FetchGranularity = 32 B for the
# random-access kernels (hashed join probe, gather from HBM) — every process below runs with BL_L2_FETCH=32 except the pytest subset
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
echo "== smoke (prebuilt library)"; SMOKE_NO_REBUILD=1 timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "group_by" -p no:cacheprovider > gpurun_out/pytest_groupby_s2b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_groupby_s2b.log | cut -c1-250
timeout 200 python tools/sweep_bulk.py > gpurun_out/r02_sweep_bulk_v6.jsonl 2> gpurun_out/sweep_bulk_v6.err; echo "sweep rc=$?"; cut -c1-230 gpurun_out/r02_sweep_bulk_v6.jsonl | tail -3
export BL_L2_FETCH=32
timeout 200 python tools/bench_kernels.py > gpurun_out/r02_kernels_l2fetch32.jsonl 2> gpurun_out/kernels_l2fetch32.err; echo "kernels rc=$?"; cut -c1-175 gpurun_out/r02_kernels_l2fetch32.jsonl
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 8 > gpurun_out/r02_bench_l2fetch32.json 2> gpurun_out/bench_l2fetch32.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_l2fetch32.json"))
print("C2", round(d["ms_per_step"], 3), d["kernels_ms_per_step"])
for s in d["secondary"]:
    print(s["config"]["workload"][60:130], round(s["ms_per_step"], 3), s["kernels_ms_per_step"])
PY
