#!/usr/bin/env python
"""Turns the raw ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under
profiles/ (run here, no GPU needed):  python scripts/summarize_profiles.py r01
  gpurun_out/launches.csv          -> profiles/<tag>_launch_list.md   (per-kernel time shares)
  gpurun_out/prof_pivot_step.ncu-rep -> profiles/<tag>_k_pivot_step_ncu.md (key metrics of the top kernel)
  gpurun_out/bench_*.json          -> profiles/<tag>_bench.md
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launch_list(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    d = collections.defaultdict(list)
    shape = {}
    for r in rows[hi + 1:]:
        if len(r) > vi:
            try:
                d[r[ki]].append(float(r[vi].replace(",", "")))
                shape[r[ki]] = (r[gi], r[bi])
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(PROF, f"{tag}_launch_list.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("Command: `ncu ... -s 3000 -c 400 python bench.py --steps 1 --warmup 1 --no-cpu` (dense 2000x2000 LP).\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | grid | block | launches | mean us | min us | max us | share of captured time |\n|---|---|---|---|---|---|---|---|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {shape[k][0]} | {shape[k][1]} | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {min(v)/1e3:.2f} | "
                    f"{max(v)/1e3:.2f} | {100*sum(v)/tot:.1f}% |\n")
    print("wrote launch list")


def ncu_full(tag):
    rep = os.path.join(OUT, "prof_pivot_step.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, f"{tag}_k_pivot_step_ncu.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of k_pivot_step (3 launches after warm-up)\n\n")
        f.write("Cache control flushes L2 before every replay, so `dram__bytes_read` shows the whole 32 MB tableau "
                "being fetched from HBM; in the real solve the tableau stays L2-resident between pivots "
                "(see the bench roofline note).\n\n| metric | unit | launch 1 | launch 2 | launch 3 |\n|---|---|---|---|---|\n")
        ki = hdr.index("Kernel Name")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in rows[2:5]) + " |\n")
        f.write("\nKernel: " + rows[2][ki] + "\n")
    print("wrote ncu summary")


def bench(tag):
    files = sorted(glob.glob(os.path.join(OUT, "bench_*.json")))
    if not files:
        return
    with open(os.path.join(PROF, f"{tag}_bench.md"), "w") as f:
        f.write(f"# {tag}: bench.py lines measured on the B200 box (gpurun)\n\n")
        for p in files:
            try:
                line = open(p).read().strip().splitlines()[-1]
                d = json.loads(line)
            except Exception:
                continue
            f.write(f"## {os.path.basename(p)}\n\n```json\n{json.dumps(d, indent=1)}\n```\n\n")
    print("wrote bench summary")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launch_list(tag)
    ncu_full(tag)
    bench(tag)
