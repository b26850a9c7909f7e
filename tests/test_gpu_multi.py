"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): node-partitioned rollout over NCCL."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_partitioned_rollout_matches_single_gpu(precision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, G4C_PRECISION=precision)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0
