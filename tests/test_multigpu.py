"""GPU, >= 2 devices: launches tests/multigpu_check.py under torchrun (one rank per GPU, NCCL) and requires its
parity verdict.  Skipped on single-GPU boxes (the gloo world_size-2 tests cover the host logic there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_partitioned_run_matches_single_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "MULTIGPU_CHECK OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
