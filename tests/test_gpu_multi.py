"""Two-rank NCCL test of the data-parallel path: spawns tests/dp_worker.py under torchrun, one
process per GPU.  Needs two GPUs (`gpurun --gpus 2`); on a one-GPU box there is nothing to exchange with, and two NCCL
ranks cannot share a device, so the test reports itself as skipped there (the world-size-2 gloo tests in
tests/test_dist_cpu.py cover the schedule and the sharding logic on the CPU)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL ranks cannot share a device)")
def test_two_rank_nccl_training_weights_sync_stats_and_sharded_fit():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0 and "dp_worker ok" in r.stdout
