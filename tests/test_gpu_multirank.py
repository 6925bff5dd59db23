"""Two ranks over NCCL (skipped with fewer than 2 GPUs): TrainStep(world_size=2) == the per-shard oracle sum.
Each case launches tests/multirank_driver.py through torch.distributed.run, one process per GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("kind,losses,port", [("ae", "autoencoder,forward,inverse", 29611), ("vae", "vae", 29612), ("dae", "dae", 29613)])
def test_two_rank_train_step_equals_per_shard_oracle(kind, losses, port):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_driver.py"), kind, losses],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, r.stdout[-4000:]
