"""GPU suite, >= 2 GPUs: the slab-decomposed path (NCCL all-to-all FFT, halo exchange, all-reduces)
against the CPU oracle. Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- pytest -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_worker(n, extra_env):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    env = dict(os.environ, CLR_TEST_N=str(n), **extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU OK" in out.stdout


@pytest.mark.parametrize("n", [64, 128])
def test_slab_decomposed_path_vs_oracle(n):
    """default transpose: peer-memory stores fused into the FFT pass (natural staging layout at these sizes)"""
    _run_worker(n, {})


@pytest.mark.parametrize("mode", ["tiled", "nccl"])
def test_slab_decomposed_path_other_transposes(mode):
    """the same checks through the tile-major staging layout of the fused transpose and through the NCCL
    all-to-all fallback (what runs when a peer's staging buffer cannot be mapped)"""
    _run_worker(64, {"COLORE_B200_P2P_TILED": "1"} if mode == "tiled" else {"COLORE_B200_P2P": "0"})


@pytest.mark.parametrize("n,tiled", [(64, 0), (128, 0), (128, 1)])
def test_slab_decomposed_path_fused_fill(n, tiled):
    """default (fp32) field kernels: the mode fill is generated inside the z pass that stores the slab transpose into
    the peers' staging buffers (fill_peer_kernel, natural and tile-major staging) -- fields against the oracle's
    Philox fill + transforms to 2e-5 sigma, everything downstream as in the other cases"""
    _run_worker(n, {"CLR_TEST_EXACT": "0", "COLORE_B200_P2P_TILED": str(tiled)})
