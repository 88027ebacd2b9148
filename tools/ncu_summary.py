"""Summarise ncu output for profiles/.
    python tools/ncu_summary.py launches gpurun_out/launches.csv            -> per-kernel launch count / time / share
    python tools/ncu_summary.py full gpurun_out/x.ncu-rep [regex]           -> key metrics per profiled launch
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = re.sub(r"^void ", "", r[ki].split("(")[0]).replace("<unnamed>::", "")[:80]
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{n}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |")
    print(f"| **all** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.1f} | 100 % |")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max"]


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    cols = {w: h.index(w) for w in WANT if w in h}
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[ki].split("(")[0]).replace("<unnamed>::", "")
        if pattern and not re.search(pattern, name):
            continue
        print(f"### `{name}`  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}")
        for w, i in cols.items():
            print(f"- {w} = {r[i]} {units[i]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
