#!/usr/bin/env python
"""Times the CPU oracle (C restatement of the reference, 1 thread) on BASELINE config 5 capped at CAP committed
nodes -- the side-by-side number of bench.py's MIP block.  Minutes of CPU: run once, result stored under profiles/."""
import json, os, sys, time, platform
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jslpsolver_b200 import problems
from oracle import ref_model

cap = int(os.environ.get("CAP", "1000"))
model = problems.knapsack_mip_model(1024, 512, seed=12345)
t0 = time.perf_counter()
sol = ref_model.solve_full(model, fast_cycles=True, max_nodes=cap, node_log=1 << 16)
dt = time.perf_counter() - t0
st = sol.state
nl = sol.tableau.node_log()
root_pivots = int(nl[0][7])
out = {"workload": f"knapsack 1024 binaries x 512 constraints, seed 12345, first {cap} committed nodes",
       "impl": "oracle/ C restatement of the reference, 1 thread, fast cycle check", "seconds": round(dt, 2),
       "iterations": int(st.bncIterations), "pivots": int(st.totalPivots), "root_pivots": root_pivots,
       "node_lps_per_s": round(st.bncIterations / dt, 3), "pivots_per_s": round(st.totalPivots / dt, 1),
       "evaluation": st.evaluation, "host": platform.processor() or platform.machine(), "cores_used": 1}
print(json.dumps(out))
with open(os.path.join(ROOT, "profiles", f"r02_cpu_config5_cap{cap}.json"), "w") as f:
    json.dump(out, f, indent=1)
