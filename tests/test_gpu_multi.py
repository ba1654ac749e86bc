"""Multi-GPU path on real GPUs (skipped on a single-GPU box; the world-size-2 gloo tests in test_host.py cover the host logic
on CPU): batch-sharded loop + the in-stream NCCL all-gather of the C ABI == torch.distributed gather == unsharded result."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_in_stream_gather_through_the_c_abi_equals_single_gpu_result():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "helpers", "gather_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
