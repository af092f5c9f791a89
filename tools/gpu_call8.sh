#!/bin/bash
# round 2, call 8 (1 GPU): plugin tests, default bench line (saved), launch list, ncu captures summarised ON THE BOX (the raw report of
# every kernel is > 64 MB and would block the copy-back), a small report with source for the four headline kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_c8.txt 2>&1; echo rc=$? >> gpurun_out/pytest_gpu_c8.txt)
tail -6 gpurun_out/pytest_gpu_c8.txt | cut -c1-300
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all_v3.json 2> gpurun_out/bench_all_v3.err; echo "bench rc=$?")
echo "== launch list of the default bench (kernel shares)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --e2e-steps 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r02_launches_bench.csv
echo "== ncu --set full, every kernel once (no source), summarised here"
timeout 1500 ncu --set full --clock-control none -k regex:"k_gb_consume|k_gbr_|k_join_probe|k_join_emit|k_join_build|k_jc_build|k_join_dense|k_part_scatter|k_gather|k_compact|k_compare|k_arith|k_gb_export|k_gb_merge_window|k_rs_scatter|k_rs_hist|k_seg_agg|k_gb_extract|k_gb_lookup" -c 60 -o /tmp/r02_all -f python tools/ncu_all.py > gpurun_out/ncu_all.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_all.log | cut -c1-200
python tools/ncu_summary.py /tmp/r02_all.ncu-rep "Round 2: ncu --set full of every product kernel (tools/ncu_all.py, 1e8-row inputs)" > gpurun_out/r02_ncu_all_kernels.md 2> gpurun_out/ncu_summary.err; wc -c gpurun_out/r02_ncu_all_kernels.md
ncu -i /tmp/r02_all.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active')]
w = csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in keep])
" > gpurun_out/r02_ncu_all_kernels_key_metrics.csv; wc -l gpurun_out/r02_ncu_all_kernels_key_metrics.csv
echo "== small report with source: the four headline kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gb_consume|k_join_probe_emit|k_gbr_scatter|k_gbr_agg" -c 6 -o gpurun_out/r02_headline -f python tools/ncu_all.py > gpurun_out/ncu_headline.log 2>&1; echo "ncu headline rc=$?"; ls -la gpurun_out/r02_headline.ncu-rep
du -sm gpurun_out
