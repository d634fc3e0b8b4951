"""Data-parallel training path over NCCL: self-launches tests/dist_nccl_check.py on 2 ranks when the box shows at least two GPUs
(the 1-GPU round-end box skips it; run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist_nccl.py -m gpu`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_training_step_over_nccl_two_ranks():
    """(1) sharded world-model loss == full-batch loss, (2) parameters bit-identical across ranks after a step and equal to the
    single-process full-batch step, (3) rollout shards concatenate to the full batch, (4) the steps replayed as CUDA graphs with
    the NCCL all-reduces inside the capture keep the ranks identical."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "tests", "dist_nccl_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "dist_nccl_check ok" in p.stdout
