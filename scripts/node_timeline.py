#!/usr/bin/env python
"""Per-CTA timeline of the resident node kernel on LargeFarmMIP (JSLP_DEBUG=1 makes the library print
one line per evaluated node to stderr).  Debug aid; run under gpurun."""
import gzip, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("NODE_TL", "1") == "1":
    os.environ["JSLP_DEBUG"] = "1"
import jslpsolver_b200 as J
from helpers import strip_timeouts
b = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "sanity_fixtures.json.gz")).read().decode())
farm = strip_timeouts([f for f in b["fixtures"] if f["file"] == "LargeFarmMIP.json"][0]["model"])
import time
import torch
spin = torch.zeros(1 << 26, device="cuda")
for rep in range(2):
    inst = J.Model().loadJson(farm)
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.5:
        spin.add_(1.0)
        torch.cuda.synchronize()
    inst.tableau.max_spec_batch = 16
    sol = inst.solve()
    print("rep", rep, sol.evaluation, inst.tableau.lastBnbStatus.gpu_ms, file=sys.stderr)
