#!/usr/bin/env python
"""us/pivot of the fused step per kernel variant on the two LP shapes that matter: BASELINE config 3 (dense
2000x2000, streaming-bound) and the config-5 root relaxation (1537x1025, selector-chain bound).  Run under gpurun.
  VARIANTS=1,10,11 SHAPES=dense2000,knaproot python scripts/variant_bench.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jslpsolver_b200 as J
from jslpsolver_b200 import _lib, problems
from jslpsolver_b200.tableau import GpuTableau

variants = [int(v) for v in os.environ.get("VARIANTS", "1,10,11").split(",")]
shapes = os.environ.get("SHAPES", "dense2000,knaproot").split(",")
for shape in shapes:
    if shape.startswith("dense"):
        n = int(shape[5:])
        it = problems.dense_packing_lp_tableau(n, n, 12345)
        g = GpuTableau(1e-8)
        g.upload(it.matrix, it.varIndexByRow, it.varIndexByCol)
    else:
        inst = J.Model().loadJson(problems.knapsack_mip_model(1024, 512, seed=12345))
        g = inst.tableau
        g.setModel(inst)
    g.save()
    ref = None
    for v in variants + variants[:1]:
        g.set_option(_lib.OPT_ENGINE, 2)
        g.set_option(_lib.OPT_STEP_VARIANT, v)
        g.set_option(_lib.OPT_GRID_PER_SM, int(os.environ.get("GRID_PER_SM", "0")))
        g.restore()
        g.simplex()
        st = g.lastStatus
        piv = st.phase1_pivots + st.phase2_pivots
        key = (piv, st.evaluation_raw)
        ref = ref or key
        print(json.dumps({"shape": shape, "variant": v, "pivots": piv, "gpu_ms": round(st.gpu_ms, 2),
                          "us_per_pivot": round(1e3 * st.gpu_ms / max(1, piv), 3), "grid_per_sm": os.environ.get("GRID_PER_SM", "0"), "same_as_first": key == ref}), flush=True)
    g.close()
