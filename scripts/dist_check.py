#!/usr/bin/env python
"""Multi-GPU parity check of the sharded branch-and-cut frontier (run with torchrun, N >= 2):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py

Every rank solves the same MIP fixtures with each round's nodes dealt round-robin over the ranks
(summaries all-gathered over NCCL) and checks, on every rank, that the committed node sequence, the
final tableau and the result are bit-identical to the oracle's sequential reference-order solve.
Rank 0 prints one line per fixture and a final DIST_CHECK OK / FAILED."""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # DIST_ONE_GPU=1: every rank on cuda:0 with the gloo backend -- exercises the sharded frontier
    # logic on a single-GPU box (NCCL needs one device per rank)
    one_gpu = os.environ.get("DIST_ONE_GPU", "0") == "1"
    if one_gpu:
        local = 0
        os.environ["JSLP_DEVICE"] = "0"
    torch.cuda.set_device(local)
    if world > 1:
        if one_gpu:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import jslpsolver_b200 as J
    from helpers import strip_timeouts
    from oracle import ref_model

    with gzip.open(os.path.join(ROOT, "tests", "golden", "sanity_fixtures.json.gz"), "rb") as f:
        bundle = json.loads(f.read().decode())
    names = ["Knapsack 1.json", "Sudoku4x4.json", "LargeFarmMIP.json", "Integer Wood Shop Problem.json",
             "Cutting stock.json", "Monster_II.json"]
    ok_all = True
    for fx in bundle["fixtures"]:
        if fx["file"] not in names:
            continue
        jm = strip_timeouts(fx["model"])
        osol = ref_model.solve_full(jm, fast_cycles=True, node_log=1 << 20)
        for K in (4, 16):
            inst = J.Model().loadJson(jm)
            inst.tableau.distributed = world > 1
            inst.tableau.shard_policy = 1  # small fixtures fit shared memory: shard them anyway, this is the check
            inst.tableau.max_spec_batch = K
            sol = inst.solve()
            gt = inst.tableau
            onl, gnl = osol.tableau.node_log(), gt.node_log()
            same = gnl.shape == onl.shape and bool(np.all((gnl == onl) | (np.isnan(gnl) & np.isnan(onl)) |
                                                          (np.arange(8)[None, :] == 3) & (onl[:, 2:3] == 0)))
            same = same and gt.branchAndCutIterations == osol.state.bncIterations
            same = same and bool(np.array_equal(gt.matrix2d(), osol.tableau.matrix()))
            same = same and sol.evaluation == osol.evaluation
            flag = torch.tensor([1 if same else 0], device="cpu" if one_gpu else "cuda")
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
            ok_all = ok_all and ok
            if rank == 0:
                b = gt.lastBnbStatus
                print(json.dumps({"fixture": fx["file"], "n_gpus": world, "spec_width": K, "ok_on_all_ranks": ok,
                                  "iterations": b.iterations, "rounds": b.rounds, "node_lps_this_rank": b.nodes_evaluated,
                                  "result": sol.evaluation}), flush=True)
    # BASELINE config 5 (knapsack 1024 x 512): node LPs in HBM slots, sharded, incumbent bound all-reduced per poll
    # (NCCL inside the library when the backend is nccl) -- against the oracle's cached first 64 nodes
    if os.environ.get("DIST_SKIP_KNAPSACK", "0") != "1":
        from jslpsolver_b200 import problems
        z = np.load(os.path.join(ROOT, "tests", "golden", "config5_knapsack.npz"))
        model = problems.knapsack_mip_model(1024, 512, seed=12345)
        for K in (8, 32):
            inst = J.Model().loadJson(model)
            inst.max_nodes = int(z["max_nodes"])
            inst.tableau.distributed = world > 1
            inst.tableau.max_spec_batch = K
            sol = inst.solve()
            gt = inst.tableau
            gnl, onl = gt.node_log(), z["node_log"]
            same = gnl.shape == onl.shape and bool(np.all((gnl == onl) | (np.arange(8)[None, :] == 3) & (onl[:, 2:3] == 0)))
            import hashlib
            same = same and hashlib.sha256(np.ascontiguousarray(gt.matrix2d()).tobytes()).hexdigest() == str(z["final_matrix_sha"])
            flag = torch.tensor([1 if same else 0], device="cpu" if one_gpu else "cuda")
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
            ok_all = ok_all and ok
            if rank == 0:
                b = gt.lastBnbStatus
                print(json.dumps({"fixture": "config 5 knapsack, first 64 nodes", "n_gpus": world, "spec_width": K,
                                  "ok_on_all_ranks": ok, "iterations": b.iterations, "rounds": b.rounds,
                                  "node_lps_this_rank": b.nodes_evaluated, "collectives": b.collectives,
                                  "gpu_ms": b.gpu_ms, "root_ms": b.host_root_ms}), flush=True)
            gt.close()
    if rank == 0:
        print("DIST_CHECK OK" if ok_all else "DIST_CHECK FAILED", flush=True)
    if world > 1:
        from jslpsolver_b200 import distributed as D
        D.destroy_communicators()
        dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
