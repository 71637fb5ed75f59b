#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and launches.csv into the small tracked summaries under profiles/.

    python profiles/summarize.py r01      # writes profiles/r01_ncu_kernels.csv, r01_launches.txt
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second",
]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    reps = sorted(f for f in os.listdir(OUT) if f.endswith(".ncu-rep"))
    with open(os.path.join(ROOT, "profiles", tag + "_ncu_kernels.csv"), "w", newline="") as fh:
        w = None
        for rep in reps:
            hdr, units, rows = raw_rows(os.path.join(OUT, rep))
            idx = [hdr.index(k) for k in KEEP if k in hdr]
            if w is None:
                w = csv.writer(fh)
                w.writerow(["report"] + ["%s [%s]" % (hdr[i], units[i]) for i in idx])
            for r in rows:
                w.writerow([rep] + [r[i] for i in idx])
    # per-launch DRAM traffic of the hot kernels, read back by bench.py for roofline.traffic
    import json
    traffic = {}
    for rep in reps:
        hdr, units, rows = raw_rows(os.path.join(OUT, rep))
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows:
            name = r[ki].split("(")[0].replace("void ", "") + "@" + rep.replace(".ncu-rep", "")
            tot = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
            traffic.setdefault(name, []).append(tot)
    with open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w") as fh:
        json.dump({k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)} for k, v in traffic.items()}, fh, indent=1, sort_keys=True)
        fh.write("\n")
    lp = os.path.join(OUT, "launches.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            v = float(r[vi].replace(",", ""))
            v = v * 1000 if r[ui] == "ms" else v / 1000 if r[ui] == "ns" else v
            a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(ROOT, "profiles", tag + "_launches.txt"), "w") as fh:
            fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 1 --skip-cpu\n")
            fh.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
            for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                fh.write("%-60s launches=%4d total_us=%10.1f share=%5.1f%% avg_us=%9.1f\n" % (k[:60], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
    print("wrote profiles/%s_*" % tag)


if __name__ == "__main__":
    main()
