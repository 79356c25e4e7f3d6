#!/usr/bin/env python
"""Root-LP timing of the knapsack MIP relaxation (phase-1 heavy, 1537x1025) on the HBM path with the
ping-pong step on/off; prints pivots per phase and us/pivot.  Run under gpurun."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import jslpsolver_b200 as J
from jslpsolver_b200 import _lib, problems
from jslpsolver_b200.tableau import GpuTableau, default_context

nk, mk = int(os.environ.get("KNAP_ITEMS", "1024")), int(os.environ.get("KNAP_CONS", "512"))
model = problems.knapsack_mip_model(nk, mk, seed=12345)
inst = J.Model().loadJson(model)

g = inst.tableau
g.setModel(inst)
g.save()
for pp in (1, 0, 1):
    g.set_option(_lib.OPT_ENGINE, 2)
    g.set_option(_lib.OPT_PINGPONG, pp)
    g.restore()
    g.simplex()
    st = g.lastStatus
    piv = st.phase1_pivots + st.phase2_pivots
    print(json.dumps({"pingpong": pp, "p1": st.phase1_pivots, "p2": st.phase2_pivots, "gpu_ms": st.gpu_ms,
                      "us_per_pivot": 1e3 * st.gpu_ms / max(1, piv), "launches": st.kernel_launches,
                      "eval": st.evaluation}), flush=True)
