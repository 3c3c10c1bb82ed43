"""The NCCL-sharded descriptor database of liborb_b200.so on >= 2 GPUs of one box (skipped on a single-GPU box; bench.py --gpus N
exercises the same entry points at N = 2 / 4 / 8)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_db_matches_unsharded():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    worker = os.path.join(os.path.dirname(__file__), "sharded_db_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29531", worker], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_DB_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
