#!/bin/bash
# round 2, session 2, final evidence run (1 GPU): full GPU test suite, the default bench line, the ncu launch list of the bench and
# one `ncu --set full` capture of every product kernel family (pages exported on the box: the report embeds the module's cubin per
# result and exceeds the 64 MiB the harness copies back).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r02_pytest_gpu_s2.txt 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r02_pytest_gpu_s2.txt | cut -c1-250
echo "== bulk sweep"; timeout 200 python tools/sweep_bulk.py > gpurun_out/r02_sweep_bulk_v5.jsonl 2> gpurun_out/sweep_bulk_v5.err; echo "sweep rc=$?"; cut -c1-230 gpurun_out/r02_sweep_bulk_v5.jsonl
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_all_v3.json 2> gpurun_out/r02_bench_all_v3.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_all_v3.json; tail -3 gpurun_out/r02_bench_all_v3.err
echo "== variants"; timeout 300 python tools/bench_variants.py > gpurun_out/r02_variants_s2.jsonl 2> gpurun_out/r02_variants_s2.err; echo "variants rc=$?"; cut -c1-260 gpurun_out/r02_variants_s2.jsonl
echo "== ncu launch list of the bench"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-verify > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_bench.csv
echo "== ncu --set full, every kernel family"
KREGEX='regex:k_gb_consume|k_gb_extract|k_gb_finalize|k_gb_export_p2p_async|k_gb_merge_window|k_gbr_hist|k_gbr_scatter|k_gbr_agg|k_seg_agg|k_join_probe|k_join_emit|k_join_build|k_join_dense_build|k_jc_|k_part_scatter|k_gather|k_compact|k_compare|k_arith|k_rs_scatter|k_str_'
timeout 900 ncu --set full --clock-control none --kernel-name "$KREGEX" -o /tmp/r02_all -f python tools/ncu_all.py > gpurun_out/ncu_all.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_all.log
ls -la /tmp/r02_all.ncu-rep
ncu -i /tmp/r02_all.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_all_raw.csv.gz
ncu -i /tmp/r02_all.ncu-rep --page details 2>/dev/null | gzip > gpurun_out/r02_ncu_all_details.txt.gz
ncu -i /tmp/r02_all.ncu-rep --page source --csv --print-source sass --kernel-name regex:k_gb_consume_lean 2>/dev/null | gzip > gpurun_out/r02_ncu_k5_lean_sass.csv.gz
python tools/ncu_summary.py /tmp/r02_all.ncu-rep "round 2 (session 2): one launch of every product kernel family, ncu --set full" > gpurun_out/r02_ncu_kernels.md 2>/dev/null; wc -l gpurun_out/r02_ncu_kernels.md
du -sm gpurun_out
