"""The sharded CUDA path under a real multi-process launch (one rank per GPU, NCCL): parity against the unsharded
single-GPU result and the oracle.  Needs >= 2 GPUs (skipped on a 1-GPU box; `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multirank.py -m gpu` runs it).  The host-side sharding arithmetic is covered on CPU by
tests/test_sharding_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_matches_unsharded_and_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multirank_worker.py"), "trafalgar-257"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTIRANK-OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
