"""Multi-GPU data parallelism on real GPUs (skipped with fewer than 2): see tests/dp_check.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_dp2_equals_single_gpu_on_concatenated_batch(mode):
    """mode "peer": the fused reduce-scatter + AdamW + all-gather kernel over NVLink peer memory (default);
    mode "nccl": NCCL all-reduce + replicated AdamW."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531" if mode == "peer" else "29532", os.path.join(root, "tests", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, DGPT_DP_MODE=mode))
    assert out.returncode == 0 and "DP_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"mode={mode}" in out.stdout
