"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: the all-gather hook that the
branch-and-cut frontier calls once per speculative round (jslp_bnb_opts.all_gather), and the
reference arm's rank discipline in bench.py."""
import ctypes as C
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from jslpsolver_b200 import distributed as D
    assert D.is_active() and D.rank_and_world() == (rank, world)
    # under gloo there is no in-library NCCL communicator: the frontier falls back to the host hook below
    assert D.nccl_communicator(object()) is None
    cb, keep = D.make_all_gather_hook()
    ok = True
    for rnd in range(3):  # several rounds with different payload sizes, like successive B&B rounds
        n = per * (rnd + 1)
        buf = np.zeros(world * n, dtype=np.uint8)
        buf[rank * n:(rank + 1) * n] = (np.arange(n) * (rank + 3) + rnd) % 251
        rc = cb(None, buf.ctypes.data, n)
        ok = ok and rc == 0
        for r in range(world):
            expect = (np.arange(n) * (r + 3) + rnd) % 251
            ok = ok and np.array_equal(buf[r * n:(r + 1) * n], expect.astype(np.uint8))
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"ok": bool(ok)}, f)
    dist.destroy_process_group()


def test_all_gather_hook_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 128, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert json.load(open(tmp_path / f"rank{r}.json"))["ok"]


def test_wire_record_layout_matches_header():
    """The all-gathered record is 16 doubles (128 bytes) per node: keep doc and code in sync."""
    src = open(os.path.join(ROOT, "jslpsolver_b200", "csrc", "jslp_frontier.h")).read()
    assert "WIRE_DOUBLES = 16" in src


def test_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under a 2-rank launch: rank 0 prints one JSON line, rank 1 exits 0."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "0", "--size", "60", "--cpu-pivots", "50"],
                           env=e, capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout.strip())
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert outs[1] == ""
