"""Per-kernel table from an `ncu --metrics ... --csv` log of several metrics (one CSV row per launch and metric):
launches, total time, DRAM bytes per launch, achieved DRAM GB/s against the measured HBM peak, L2->SM bytes, tensor-pipe
and issue-slot activity.  Used for profiles/r2_summary.md (the kernels that are NOT the GEMM).
    python tools/ncu_metrics_table.py gpurun_out/r2_train_metrics.csv [hbm_gbs]"""
import collections
import csv
import json
import os
import re
import sys


def main():
    path = sys.argv[1]
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm = float(sys.argv[2]) if len(sys.argv) > 2 else (json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0)
    rows = list(csv.reader(open(path, errors="replace")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.defaultdict(dict)          # launch id -> metric -> value
    names = {}
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        try:
            per[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        names[r[ii]] = re.sub(r"^void ", "", r[ki].split("(")[0]).replace("<unnamed>::", "")[:64]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for lid, m in per.items():
        a = agg[names[lid]]
        a["n"] += 1
        t = m.get("gpu__time_duration.sum", 0.0)
        a["ns"] += t
        a["rd"] += m.get("dram__bytes_read.sum", 0.0)
        a["wr"] += m.get("dram__bytes_write.sum", 0.0)
        a["l2"] += m.get("lts__t_bytes.sum", 0.0)
        a["tensor_w"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
        a["issue_w"] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * t
    tot = sum(a["ns"] for a in agg.values())
    print("| kernel | launches | total us | us / launch | DRAM MB / launch (rd + wr) | DRAM GB/s | of HBM peak %.0f | L2->SM MB / launch | "
          "tensor pipe %% | issue slots %% |" % hbm)
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        c = a["n"]
        gbs = (a["rd"] + a["wr"]) / a["ns"] if a["ns"] else 0.0          # bytes / ns = GB/s
        print(f"| `{n}` | {int(c)} | {a['ns'] / 1e3:.1f} | {a['ns'] / c / 1e3:.2f} | {a['rd'] / c / 1e6:.2f} + {a['wr'] / c / 1e6:.2f} | "
              f"{gbs:.0f} | {gbs / hbm:.2f} | {a['l2'] / c / 1e6:.1f} | {a['tensor_w'] / a['ns'] if a['ns'] else 0:.1f} | "
              f"{a['issue_w'] / a['ns'] if a['ns'] else 0:.1f} |")
    print(f"| **all** | {int(sum(a['n'] for a in agg.values()))} | {tot / 1e3:.1f} | | | | | | | |")


if __name__ == "__main__":
    main()
