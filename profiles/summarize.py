#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and launches.csv into the small tracked summaries under profiles/.

    python profiles/summarize.py r02      # writes profiles/r02_ncu_kernels.csv, r02_traffic.json, r02_launches.txt
"""
import collections
import csv
import io
import json
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
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second",
]
# ncu picks a unit per REPORT (us in one, ms in the next): every value is converted to one fixed unit per
# column family, named in the header, so rows of different reports can share a file
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}     # -> us
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}                                              # -> byte


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def norm(value, unit):
    """-> (value in the column's fixed unit, that unit)"""
    try:
        v = float(value.replace(",", ""))
    except ValueError:
        return value, unit
    if unit in TIME:
        return v * TIME[unit], "us"
    if unit in BYTES:
        return v * BYTES[unit], "byte"
    return v, unit


def git_head():
    try:
        head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
        dirty = bool(subprocess.run(["git", "status", "--porcelain", "--", "fiber_b200", "include"], cwd=ROOT, capture_output=True, text=True).stdout.strip())
        return head + ("+dirty" if dirty else "")
    except Exception:       # noqa: BLE001
        return None


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    reps = sorted(f for f in os.listdir(OUT) if f.endswith(".ncu-rep"))
    traffic = {}
    with open(os.path.join(ROOT, "profiles", tag + "_ncu_kernels.csv"), "w", newline="") as fh:
        w = None
        for rep in reps:
            hdr, units, rows = raw_rows(os.path.join(OUT, rep))
            idx = [hdr.index(k) for k in KEEP if k in hdr]
            if w is None:
                w = csv.writer(fh)
                w.writerow(["report"] + ["%s [%s]" % (hdr[i], norm("0", units[i])[1]) for i in idx])
            for r in rows:
                w.writerow([rep] + [norm(r[i], units[i])[0] for i in idx])
            # per-launch DRAM traffic / L2 load of the hot kernels, read back by bench.py for roofline.traffic
            ki = hdr.index("Kernel Name")

            def col(name, r):
                if name not in hdr:
                    return None
                i = hdr.index(name)
                v = norm(r[i], units[i])[0]
                return v if isinstance(v, float) else None
            for r in rows:
                name = r[ki].split("(")[0].replace("void ", "") + "@" + rep.replace(".ncu-rep", "")
                t = traffic.setdefault(name, {"dram": [], "lts_pct": [], "lts_bytes": [], "us": []})
                t["dram"].append((col("dram__bytes_read.sum", r) or 0.0) + (col("dram__bytes_write.sum", r) or 0.0))
                t["lts_pct"].append(col("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", r))
                t["lts_bytes"].append(col("lts__t_bytes.sum", r))
                t["us"].append(col("gpu__time_duration.sum", r))

    def mean(v):
        v = [x for x in v if x is not None]
        return sum(v) / len(v) if v else None
    summary = {k: {"dram_bytes_per_launch": mean(t["dram"]), "lts_pct_of_peak": mean(t["lts_pct"]), "lts_bytes_per_launch": mean(t["lts_bytes"]),
                   "ncu_time_us": mean(t["us"]), "launches": len(t["dram"])} for k, t in traffic.items()}
    src_sha = None
    try:
        src_sha = open(os.path.join(OUT, "src_sha256.txt")).read().strip()
    except OSError:
        pass
    summary["_meta"] = {"commit": git_head(), "src_sha256": src_sha,
                        "how": "ncu --set full --clock-control none (profiles/run_ncu.sh); caches are flushed between replays, times are cold"}
    with open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w") as fh:
        json.dump(summary, fh, indent=1, sort_keys=True)
        fh.write("\n")
    lp = os.path.join(OUT, "launches.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            v, _ = norm(r[vi], r[ui])
            a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(ROOT, "profiles", tag + "_launches.txt"), "w") as fh:
            fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 1 --skip-cpu --skip-parzen\n")
            fh.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes;  build %s\n" % summary["_meta"]["commit"])
            for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                fh.write("%-60s launches=%4d total_us=%10.1f share=%5.1f%% avg_us=%9.1f\n" % (k[:60], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
    print("wrote profiles/%s_*" % tag)


if __name__ == "__main__":
    main()
