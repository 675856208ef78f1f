"""NCCL check on >= 2 GPUs (skipped on a single-GPU box): tools/check_multigpu.py under torchrun verifies that the
user-sharded rebuild + edge all-gather equals the single-GPU result bit for bit and that the row-partitioned
propagation equals the full SpMM."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_rebuild_and_partitioned_propagation_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_multigpu.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
