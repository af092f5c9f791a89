#!/bin/bash
# usage: tools/mgpu_run2.sh N   — correctness check of every multi-GPU plan + two weak-scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== mgpu_check"; timeout 600 $TR tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "exit $?"; grep -E "mgpu_check|Error|error|assert" gpurun_out/mgpu_check_$N.log | head -20
for sk in uniform zipf; do
  echo "== bench --gpus $N --exchange p2p --skew $sk"
  timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --exchange p2p --e2e-steps 1 --skew $sk > gpurun_out/bench_g${N}_$sk.json 2> gpurun_out/bench_g${N}_$sk.err; echo "exit $?"
  python -c "import json; d=[json.loads(l) for l in open('gpurun_out/bench_g${N}_$sk.json') if l.startswith('{')][-1]; print('$sk', d['n_gpus'], round(d['ms_per_step'],3), f\"{d['value']:.3e}\", {k: round(v,3) for k,v in d['kernels_ms_per_step'].items() if v > 0.02}, 'e2e', f\"{d['e2e']['value']:.3e}\")" || tail -5 gpurun_out/bench_g${N}_$sk.err
done
