"""Multi-GPU parity inside `pytest -m gpu` (skipped on a box with fewer than 2 devices): the partitioned group_by
(NCCL all-to-all and the fused P2P-store exchange whose counts/flags travel through the peer windows), the partitioned
and the broadcast hash join and the raw-row group_by plan, each against a numpy reduction of the concatenated inputs and
against the reference's partition function (hash_to_partition(dirty_hash(key), P), polars-utils/src/hashing.rs:62-69) —
tools/mgpu_check.py, one process per GPU under torchrun."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
def test_partitioned_plans_on_two_gpus():
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + (os.getpid() % 400)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "mgpu_check.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    for tag in ("mgpu_check nccl: OK", "mgpu_check p2p: OK", "mgpu_check join: OK", "mgpu_check broadcast join: OK", "mgpu_check raw-row group_by: OK"):
        assert tag in r.stdout, (tag, r.stdout[-3000:])
