"""Multi-GPU branch-and-cut (needs >= 2 GPUs on the box; skipped otherwise): the frontier sharded
over ranks must commit exactly the node sequence of the reference's sequential loop."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_frontier_matches_oracle_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "scripts", "dist_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "DIST_CHECK OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_sharded_frontier_two_ranks_sharing_one_gpu():
    """Same check with both ranks on cuda:0 and gloo carrying the all-gather: runs on a one-GPU box."""
    env = dict(os.environ, DIST_ONE_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "scripts", "dist_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "DIST_CHECK OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
