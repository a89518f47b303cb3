"""Reads `ncu -i <rep> --page raw --csv` exports and writes / updates profiles/ncu_traffic.json: per workload the DRAM
bytes of ONE launch of the dominant kernel (the `traffic` figure of bench.py's roofline object) plus a few context
metrics.  usage: python tools/ncu_traffic.py <workload> <raw.csv> [kernel-substring]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "gpu__time_duration.sum": "duration", "launch__registers_per_thread": "registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_slots_pct"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
        "nsecond": 1e-9, "second": 1.0}


def main():
    workload, path = sys.argv[1], sys.argv[2]
    pick = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    out = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names) or (pick and pick not in r[col["Kernel Name"]]):
            continue
        e = {"kernel": r[col["Kernel Name"]][:120]}
        for m, key in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", ""))
                e[key] = v * UNIT.get(units[col[m]], 1.0)
        out.append(e)
    if not out:
        raise SystemExit("no matching kernel rows")
    # mean over the captured launches of the kernel (mid-sweep layers / waves / passes), with the spread
    best = {"kernel": out[0]["kernel"]}
    for key in WANT.values():
        vals = [e[key] for e in out if key in e]
        if vals:
            best[key] = sum(vals) / len(vals)
    d_ = [e["duration"] for e in out if "duration" in e]
    if d_:
        best["duration_min"], best["duration_max"] = min(d_), max(d_)
    best["launches_captured"] = len(out)
    best["source"] = os.path.basename(path)
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    d[workload] = best
    json.dump(d, open(p, "w"), indent=1, sort_keys=True)
    print(workload, best)


if __name__ == "__main__":
    main()
