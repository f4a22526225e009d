"""Multi-GPU parity + stress on a real box (tools/gpu_dist_check.py under torchrun): sharded write pass — fused
peer-memory exchange (eager and as one captured CUDA graph per step), NCCL all-gather — against the unsharded
processor over 100 steps = 200 exchange epochs with one rank of every CFG half randomly late, then the story finishes
with frame-parallel reads from the all-gathered id_bank.  Needs >= 2 B200s; skipped on the single-GPU test tier."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_story_on_gpus(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "gpu_dist_check.py"),
           "--steps", "100", "--skew", "--graph"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "gpu_dist_check ok" in r.stdout
