"""GPU parity at N=2: sharded fused PAN path with NVLink peer halo rows, and the band alignment sharded by section, vs the
whole-strip oracle; the offset-estimation sums through one all-reduce."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_gpu_shards_match_oracle(oracle_mod):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-3000:] + "\n".join(l for l in r.stderr.splitlines() if "site-packages/torch/distributed" not in l)[-6000:]
