"""Data-parallel path: host contract on CPU (gloo, world_size 2) and the fused NVLink exchange on GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(world, backend, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), backend]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_dp_contract_gloo_world2():
    r = _torchrun(2, "gloo", 29611)
    assert r.returncode == 0 and "DP_GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_batch_sharding_helpers(eh):
    import numpy as np
    from easyhybrid_b200.dp import combine_mse, global_batch_indices
    perms = [np.arange(10)[::-1], np.arange(10)]
    gi = global_batch_indices(perms, [10, 10], 4, 1)
    assert gi.tolist() == [5, 4, 3, 2, 14, 15, 16, 17]
    L, g = combine_mse([2, 6], [1.0, 3.0], [[1.0, 0.0], [3.0, 4.0]])
    assert L == pytest.approx(2.5) and g.tolist() == pytest.approx([2.5, 3.0])


@pytest.mark.gpu
def test_dp_fused_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = _torchrun(2, "nccl", 29612)
    assert r.returncode == 0 and "DP_NCCL_OK" in r.stdout and "DP_NCCL_HOST_OK" in r.stdout and "DP_NCCL_STATS_OK" in r.stdout and "DP_NCCL_WIDE_OK" in r.stdout and "DP_NCCL_JIT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
