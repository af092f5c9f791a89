#!/bin/bash
# usage: tools/mgpu_run.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== mgpu_check"; timeout 600 $TR tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "exit $?"; grep -E "mgpu_check|Error|error|assert" gpurun_out/mgpu_check_$N.log | head -20
for ex in p2p nccl; do
  echo "== bench --gpus $N --exchange $ex"
  timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --exchange $ex --e2e-steps 1 > gpurun_out/bench_g${N}_$ex.json 2> gpurun_out/bench_g${N}_$ex.err; echo "exit $?"
  python -c "import json; d=json.load(open('gpurun_out/bench_g${N}_$ex.json')); print('$ex', d['n_gpus'], round(d['ms_per_step'],3), f\"{d['value']:.3e}\", {k: round(v,3) for k,v in d['kernels_ms_per_step'].items() if v > 0.02}, 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/bench_g${N}_$ex.err
done
