"""GPU suite, >= 2 GPUs: the slab-decomposed path (NCCL all-to-all FFT, halo exchange, all-reduces)
against the CPU oracle. Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- pytest -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [64, 128])
def test_slab_decomposed_path_vs_oracle(n):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    env = dict(os.environ, CLR_TEST_N=str(n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU OK" in out.stdout
