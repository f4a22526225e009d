"""Multi-GPU parity on a real box: sharded write pass (fused peer-memory exchange and NCCL all-gather) against the
unsharded processor.  Needs >= 4 B200s; skipped on the single-GPU test tier (tools/gpu_dist_check.py is what runs)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [4, 8])
def test_sharded_write_pass_on_gpus(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "gpu_dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "gpu_dist_check ok" in r.stdout
