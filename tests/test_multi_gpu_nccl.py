"""
Multi-GPU test with the REAL kernels and NCCL (needs >= 2 GPUs; run with `gpurun --gpus 2`): G-GPU results are
bit-identical to the single-GPU result for both partitions, through the gather-to-one-rank path, the broadcast
of the source mesh and the api driver.  See tests/dist_gpu_worker.py.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpus_bit_identical_with_real_kernels(tmp_path):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_worker.py"),
           str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(tmp_path / "ok")
