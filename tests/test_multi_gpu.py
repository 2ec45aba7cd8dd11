"""World size > 1 on real GPUs: sharded == single-rank, bit for bit, for every estimator family
(MI, CMI with own-order marginals, O-information, TC, entropy, KL divergence, KSG-2, a p-norm CMI,
the mixed branch, sample entropy).  Spawns ``tools/check_multi_gpu.py`` under torchrun with NCCL
when at least two devices are visible; skipped on a single-GPU box (the CPU suite covers the
host-side sharding logic with gloo, tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_single_rank_on_two_gpus():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip(f"needs >= 2 GPUs (visible: {ngpu})")
    world = 2
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:]
    assert "False" not in out.stdout
