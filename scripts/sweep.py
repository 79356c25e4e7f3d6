#!/usr/bin/env python
"""Tuning sweep of the fused pivot step on the dense 2000x2000 LP (run under gpurun).
Prints pivots/s per configuration and a per-CTA timeline breakdown; writes gpurun_out/sweep.json."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from jslpsolver_b200 import _lib, problems
from jslpsolver_b200.tableau import DeviceContext, GpuTableau


def timeline(g, clock_ghz=1.965):
    L = g.context.lib
    n, grid = C.c_int(), C.c_int()
    _lib.check(L.jslp_debug_timeline(g.handle, None, 0, C.byref(n), C.byref(grid)))
    if n.value == 0:
        return None
    buf = np.zeros((n.value, grid.value, 8), dtype=np.int64)
    _lib.check(L.jslp_debug_timeline(g.handle, buf.ctypes.data, buf.size, C.byref(n), C.byref(grid)))
    buf = buf[8:]  # skip the first launches (bootstrap)
    us = lambda cyc: cyc / (clock_ghz * 1e3)
    staged, rows, ticket, done = (us(buf[:, :, k].astype(np.float64)) for k in (1, 2, 3, 4))
    last = buf[:, :, 6] == 1
    g0 = buf[:, :, 0].astype(np.float64)
    g0 = np.where(g0 > 1e12, g0, np.nan)  # globaltimer stamps only (the selector's slot 0 holds a cycle count)
    start_spread = (np.nanmax(g0, axis=1) - np.nanmin(g0, axis=1)) / 1e3
    launch_period = np.diff(np.nanmin(g0, axis=1)) / 1e3
    sel_exit = float(done[last].mean()) if last.any() else None
    rowm = buf[:, :, 6] == 0   # row CTAs only (1 = deciding selector, 2 = staging selector)
    sel = {}
    pub = rows[rowm].reshape(buf.shape[0], -1) if rowm.sum() % buf.shape[0] == 0 else None
    norm = staged[rowm].reshape(buf.shape[0], -1) if pub is not None else None
    if last.any():  # selector record: d1 norm, d2 arrivals seen, d3 partials reduced, d0 rows derived, d4 exit
        sel = {"sel_norm": float(staged[last].mean()), "sel_arrivals_seen": float(rows[last].mean()),
               "sel_partials_reduced": float(ticket[last].mean()),
               "sel_rows_derived": float(us(buf[:, :, 0].astype(np.float64))[last].mean()),
               "sel_exit": sel_exit,
               "last_publish_us": float(np.where(rowm, rows, 0).max(axis=1).mean())}
    out = {
        "launches": int(buf.shape[0]), "grid": int(grid.value),
        "t1_norm_us": float(staged[rowm].mean()), "t2_publish_us": float(rows[rowm].mean()),
        "t3_rows_done_us": float(ticket[rowm].mean()), "t3_rows_done_max_us": float(ticket.max(axis=1).mean()),
        "selector_exit_us": sel_exit, "selector": sel,
        "stage_us_mean": float(staged.mean()), "stage_us_max": float(staged.max(axis=1).mean()),
        "rows_us_mean": float((rows - staged).mean()), "rows_us_max": float((rows - staged).max(axis=1).mean()),
        "to_ticket_us_mean": float(ticket.mean()), "to_ticket_us_max": float(ticket.max(axis=1).mean()),
        "tail_us_mean": float((done - ticket)[last].mean()) if last.any() else None,
        "cta_total_us_max": float(done.max(axis=1).mean()),
        "cta_start_spread_us": float(start_spread.mean()),
        "publish_pct_us": [float(x) for x in np.percentile(pub, [5, 50, 95, 99, 100])] if pub is not None else None,
        "publish_launch_max_us": float(pub.max(axis=1).mean()) if pub is not None else None,
        "norm_pct_us": [float(x) for x in np.percentile(norm, [5, 50, 95, 100])] if norm is not None else None,
        "launch_period_us": float(np.median(launch_period)) if len(launch_period) else None,
    }
    return out


def main():
    size = int(os.environ.get("SWEEP_SIZE", "2000"))
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = DeviceContext(0, stream.cuda_stream)
    it = problems.dense_packing_lp_tableau(size, size, 12345)
    H, W = it.matrix.shape
    g = GpuTableau(1e-8, context=ctx)
    g.upload(it.matrix, it.varIndexByRow, it.varIndexByCol, row_capacity=H)
    g.save()
    configs = []
    for variant in (0, 3, 1, 2, 4, 6):
        for pdl in (0, 1):
            configs.append({"engine": 2, "variant": variant, "grid": 0, "look": 1, "pdl": pdl})
    configs.append({"engine": 2, "variant": 0, "grid": 0, "look": 0, "pdl": 0})
    configs.append({"engine": 2, "variant": 0, "grid": 0, "look": 0, "pdl": 1})
    configs.append({"engine": 2, "variant": 0, "grid": 1, "look": 1, "pdl": 1})
    configs.append({"engine": 2, "variant": 2, "grid": 3, "look": 1, "pdl": 1})
    configs.append({"engine": 1, "variant": 0, "grid": 0, "look": 0, "pdl": 0})
    if os.environ.get("QUICK", "0") == "1":
        configs = [{"engine": 2, "variant": v, "grid": 0, "look": 1, "pdl": 0, "pp": 1} for v in [int(x) for x in os.environ.get("VARIANTS", "1,6,9").split(",")]]
    results = []
    for cfg in configs:
        g.set_option(_lib.OPT_PINGPONG, cfg.get("pp", 1))
        g.set_option(_lib.OPT_ENGINE, cfg["engine"])
        g.set_option(_lib.OPT_STEP_VARIANT, cfg["variant"])
        g.set_option(_lib.OPT_GRID_PER_SM, cfg["grid"])
        g.set_option(_lib.OPT_LOOKAHEAD, cfg["look"])
        g.set_option(_lib.OPT_PDL, cfg["pdl"])
        g.set_option(_lib.OPT_TIMELINE, 0)
        best = None
        for rep in range(3):
            g.restore()
            g.simplex()
            st = g.lastStatus
            piv = st.phase1_pivots + st.phase2_pivots
            us = 1e3 * st.gpu_ms / max(1, piv)
            best = us if best is None else min(best, us)
        r = dict(cfg)
        r.update({"pivots": piv, "us_per_pivot": best, "pivots_per_s": 1e6 / best, "eval": st.evaluation})
        # timeline of the first 200 pivots
        g.set_option(_lib.OPT_TIMELINE, 200)
        g.restore()
        g.simplex()
        r["timeline"] = timeline(g)
        results.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
