#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
timeout 170 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
