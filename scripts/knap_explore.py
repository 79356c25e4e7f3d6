#!/usr/bin/env python
"""How deep does branch and cut go on BASELINE config 5 (knapsack 1024 x 512, seed 12345) for a given
`tolerance`?  Used once to calibrate the tolerance bench.py's MIP leg runs to termination with.
  TOLS=0.02,0.01 CAP=20000 SPEC=32 python scripts/knap_explore.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import jslpsolver_b200 as J
from jslpsolver_b200 import problems

nk, mk = int(os.environ.get("KNAP_ITEMS", "1024")), int(os.environ.get("KNAP_CONS", "512"))
for tol in [float(x) for x in os.environ.get("TOLS", "0.02,0.01").split(",")]:
    model = problems.knapsack_mip_model(nk, mk, seed=12345, tolerance=tol)
    inst = J.Model().loadJson(model)
    inst.max_nodes = int(os.environ.get("CAP", "20000"))
    inst.tableau.max_spec_batch = int(os.environ.get("SPEC", "32"))
    t0 = time.perf_counter()
    sol = inst.solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    b = inst.tableau.lastBnbStatus
    print(json.dumps({"tolerance": tol, "iterations": b.iterations, "node_lps": b.nodes_evaluated, "rounds": b.rounds,
                      "pivots": b.pivots, "wall_s": round(dt, 3), "gpu_ms": round(b.gpu_ms, 1), "root_ms": round(b.host_root_ms, 1),
                      "result": sol.evaluation, "best_possible": -b.best_possible_eval, "is_integral": b.is_integral,
                      "capped": b.iterations >= inst.max_nodes}), flush=True)
    inst.tableau.close()
