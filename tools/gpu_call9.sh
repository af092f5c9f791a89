#!/bin/bash
# round 2, call 9 (1 GPU): the evidence of call 8 again, this time with a size guard on gpurun_out (the copy-back limit is 64 MiB)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_c9.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c9.txt)
tail -5 gpurun_out/pytest_gpu_c9.txt | cut -c1-300
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all_v3.json 2> gpurun_out/bench_all_v3.err; echo "bench rc=$?")
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_all_v3.json").read().strip().splitlines()[-1])
    print("C2", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), d["verified"][:30], "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in d["kernels_ms_per_step"].items()})
    for s in d.get("secondary", []): print("C3", round(s["ms_per_step"],3), s["roofline"]["kernel"], round(s["roofline"]["frac"],4), str(s["verified"])[:20], "e2e", round(s["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in s["kernels_ms_per_step"].items()})
except Exception as e: print("ERR", e, open("gpurun_out/bench_all_v3.err").read()[-800:])
PY
timeout 300 python tools/bench_kernels.py > gpurun_out/kernels_c9.jsonl 2> gpurun_out/kernels_c9.err; grep -E "k2_compare|k3_compact" gpurun_out/kernels_c9.jsonl | cut -c1-200
echo "== launch list of the default bench (kernel shares)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --e2e-steps 0 > /tmp/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r02_launches_bench.csv
echo "== ncu --set full, every kernel once (no source), summarised here; the report itself stays on the box"
timeout 1500 ncu --set full --clock-control none -k regex:"k_gb_consume|k_gbr_|k_join_probe|k_join_emit|k_join_build|k_jc_build|k_join_dense|k_part_scatter|k_gather|k_compact|k_compare|k_arith|k_gb_export|k_gb_merge_window|k_rs_scatter|k_rs_hist|k_seg_agg|k_gb_extract|k_gb_lookup" -c 60 -o /tmp/r02_all -f python tools/ncu_all.py > /tmp/ncu_all.log 2>&1; echo "ncu full rc=$?"; tail -2 /tmp/ncu_all.log | cut -c1-200
python tools/ncu_summary.py /tmp/r02_all.ncu-rep "Round 2: ncu --set full of every product kernel (tools/ncu_all.py, 1e8-row inputs)" > gpurun_out/r02_ncu_all_kernels.md 2> /tmp/ncu_summary.err; wc -c gpurun_out/r02_ncu_all_kernels.md
echo "== headline kernels with source (one launch each), source page extracted here"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gb_consume" -s 3 -c 1 -o /tmp/r02_k5 -f python bench.py --workload groupby --steps 2 --warmup 1 --no-cpu-baseline --no-verify --e2e-steps 0 > /tmp/ncu_k5.log 2>&1; echo "ncu k5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_join_probe_emit" -s 3 -c 1 -o /tmp/r02_k8 -f python bench.py --workload join --join-keys sparse --steps 2 --warmup 1 --no-cpu-baseline --no-verify --e2e-steps 0 > /tmp/ncu_k8.log 2>&1; echo "ncu k8 rc=$?"
for k in k5 k8; do
  ls -la /tmp/r02_$k.ncu-rep
  ncu -i /tmp/r02_$k.ncu-rep --page source --csv > gpurun_out/r02_ncu_${k}_source.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/r02_$k.ncu-rep "Round 2: $k headline kernel inside bench.py" > gpurun_out/r02_ncu_${k}_headline.md 2>/dev/null
  sz=$(stat -c %s /tmp/r02_$k.ncu-rep); if [ "$sz" -lt 20000000 ]; then cp /tmp/r02_$k.ncu-rep gpurun_out/; fi
done
du -sm gpurun_out; ls -la gpurun_out | awk '{print $5, $9}' | sort -rn | head -5
total=$(du -sm gpurun_out | cut -f1); if [ "$total" -gt 55 ]; then rm -f gpurun_out/*.ncu-rep; echo "reports dropped to stay under the copy-back limit"; du -sm gpurun_out; fi
