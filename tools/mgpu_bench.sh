#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for ex in p2p nccl; do
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --exchange $ex --e2e-steps 1 > gpurun_out/bench_g${N}_$ex.json 2> gpurun_out/bench_g${N}_$ex.err; echo "exit $?"
  python -c "import json; d=[json.loads(l) for l in open('gpurun_out/bench_g${N}_$ex.json') if l.startswith('{')][-1]; print('$ex', d['n_gpus'], round(d['ms_per_step'],3), f\"{d['value']:.3e}\", {k: round(v,3) for k,v in d['kernels_ms_per_step'].items() if v > 0.02}, 'e2e', f\"{d['e2e']['value']:.3e}\")" || tail -5 gpurun_out/bench_g${N}_$ex.err
  head -c 200 gpurun_out/bench_g${N}_$ex.json | head -2 | cut -c1-80
done
