"""Multi-GPU GPU tests (need >= 2 visible GPUs; skipped on a single-GPU box): the fused encoder -> all-gather kernel
(tcgen05 GEMM epilogue storing into every rank's symmetric-memory buffer over NVLink) against encoder + NCCL
all_gather_into_tensor, including the gathered PLN loss and its embedding gradient - all bit-identical."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("multicast", ["1", "0"])
def test_fused_encoder_gather_two_gpus(multicast):
    env = dict(os.environ, OSR_MC=multicast)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541" if multicast == "1" else "29542",
           os.path.join(ROOT, "tools", "symm_probe", "fused_gather_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "bit-identical to encoder + NCCL all-gather (incl. loss, grads): True" in out.stdout, out.stdout[-2000:]
