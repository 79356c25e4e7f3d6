#!/usr/bin/env python
"""Summarises an `ncu --page raw --csv` dump of k_pivot_step launches into the JSON bench.py reads for
`roofline.traffic` (profiles/r02_k_pivot_step_ncu.json) and a markdown table for profiles/.

  ncu --set full --cache-control none --clock-control none -k regex:k_pivot_step -s 3000 -c 6 -o gpurun_out/pp python bench.py ...
  ncu -i gpurun_out/pp.ncu-rep --page raw --csv > gpurun_out/pp_raw.csv
  python scripts/ncu_summary.py gpurun_out/pp_raw.csv profiles/r02_k_pivot_step_ncu "ncu --set full --cache-control none ..."
"""
import csv, json, sys

WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "lts__t_bytes.sum": "lts_bytes",
        "gpu__time_duration.sum": "duration", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed": "mem_pct",
        "launch__grid_size": "grid", "launch__registers_per_thread": "regs",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3,
        "msecond": 1e3, "ms": 1e3, "second": 1e6}


def main():
    src, out, how = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = [r for r in csv.reader(open(src)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    col = {n: i for i, n in enumerate(names)}
    launches = []
    for r in data:
        if len(r) != len(names):
            continue
        rec = {"kernel": r[col["Kernel Name"]][:60]}
        for k, short in WANT.items():
            if k in col:
                try:
                    rec[short] = float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)
                except ValueError:
                    pass
        launches.append(rec)
    if not launches:
        raise SystemExit("no launches parsed")
    avg = lambda k: sum(l.get(k, 0.0) for l in launches) / len(launches)
    summary = {"source": f"{out}.md ({how})", "launches": len(launches),
               "dram_bytes_per_launch": int(avg("dram_read") + avg("dram_write")),
               "dram_read_per_launch": int(avg("dram_read")), "dram_write_per_launch": int(avg("dram_write")),
               "lts_bytes_per_launch": int(avg("lts_bytes")), "duration_us": avg("duration"),
               "l2_hit_pct": avg("l2_hit_pct"), "dram_pct_of_peak": avg("dram_pct"), "lts_pct_of_peak": avg("lts_pct")}
    json.dump(summary, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# k_pivot_step, warm cache, mid-solve launches\n\n`{how}`\n\n")
        f.write("| launch | grid | regs | duration us | DRAM read MB | DRAM write MB | L2 (lts) MB | L2 hit % | DRAM % peak | L2 % peak |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for i, l in enumerate(launches):
            f.write(f"| {i} | {l.get('grid', 0):.0f} | {l.get('regs', 0):.0f} | {l.get('duration', 0):.2f} | {l.get('dram_read', 0) / 1e6:.2f} | "
                    f"{l.get('dram_write', 0) / 1e6:.2f} | {l.get('lts_bytes', 0) / 1e6:.1f} | {l.get('l2_hit_pct', 0):.1f} | "
                    f"{l.get('dram_pct', 0):.1f} | {l.get('lts_pct', 0):.1f} |\n")
        f.write("\n```json\n" + json.dumps(summary, indent=1) + "\n```\n")
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
