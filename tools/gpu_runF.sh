#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi or tuples or heavy or kats" -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_multi.log | cut -c1-400
