"""rank r's result under torchrun (NCCL weight broadcast from rank 0, one sample per
GPU, seed 42 + r) is bit-identical to what a single process produces for that seed.  Needs >= 2 GPUs (skipped otherwise;
run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_rank_results_equal_single_gpu_results(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 4)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(here, "rank_identity_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    single = torch.load(tmp_path / "single.pt")
    for rank in range(world):
        mine = torch.load(tmp_path / f"rank{rank}.pt")
        assert torch.isfinite(mine).all() and torch.equal(mine, single[rank]), rank
    assert not torch.equal(single[0], single[1])  # different seeds, different samples
