#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 0 --steps 5 --warmup 3"
echo "== key-load flavour"
for v in "BL_K5_HINT=0" "BL_K5_HINT=2" "BL_K5_HINT=3"; do
  env $v timeout 300 $B 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['knobs'], round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4))"
done
timeout 300 $B --keys 1500000 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('keys=1.5M', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if v>0.05})"
BL_K5_HINT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "group_by" 2>&1 | tail -2
echo "== compute-sanitizer memcheck on a test subset"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin_abi.py -m gpu -q -p no:cacheprovider -x -k "kats or smem_plan or streaming or edge or dtypes or chunked or sliced or error or null_scalar or plugin or gather or filter_kat" > gpurun_out/sanitizer_tests.log 2>&1; echo "sanitizer exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitizer_tests.log | head -8
