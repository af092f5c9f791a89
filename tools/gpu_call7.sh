#!/bin/bash
# round 2, call 7 (1 GPU): parity after table_hash, default bench line, ncu captures of every kernel, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_c7.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c7.txt)
tail -6 gpurun_out/pytest_gpu_c7.txt | cut -c1-300
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all_v2.json 2> gpurun_out/bench_all_v2.err; echo "bench rc=$?")
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_all_v2.json").read().strip().splitlines()[-1])
    print("C2", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), d["verified"][:30], "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in d["kernels_ms_per_step"].items()})
    for s in d.get("secondary", []): print("C3", round(s["ms_per_step"],3), s["roofline"]["kernel"], round(s["roofline"]["frac"],4), str(s["verified"])[:20], "e2e", round(s["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in s["kernels_ms_per_step"].items()})
    cb = d.get("cpu_baseline", {}); print("cpu", [(p["cores"], f'{p["value"]:.2e}') for p in cb.get("thread_scaling", [])], [(p["cores"], f'{p["value"]:.2e}') for p in cb.get("join", {}).get("thread_scaling", [])])
except Exception as e: print("ERR", e, open("gpurun_out/bench_all_v2.err").read()[-800:])
PY
echo "== launch list of the default bench (kernel shares)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --e2e-steps 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r02_launches_bench.csv
echo "== ncu --set full, every kernel once"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_gb_consume|k_gbr_|k_join_probe|k_join_emit|k_join_build|k_jc_build|k_join_dense|k_part_scatter|k_gather|k_compact|k_compare|k_arith|k_gb_export|k_gb_merge_window|k_rs_scatter|k_rs_hist|k_seg_agg|k_gb_extract|k_gb_lookup" -c 60 -o gpurun_out/r02_all -f python tools/ncu_all.py > gpurun_out/ncu_all.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_all.log | cut -c1-300; ls -la gpurun_out/r02_all.ncu-rep
