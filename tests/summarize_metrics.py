"""Pivot an `ncu --metrics a,b,c --csv` log (one row per launch and metric) into one line per launch.

    python tests/summarize_metrics.py gpurun_out/metrics.csv [tiles]

Not a pytest module."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 74
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rows = OrderedDict()
for r in csv.DictReader(lines):
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = re.sub(r"\(.*", "", name).replace("unnamed>::", "").replace("fb::", "").replace("<", "", 1)
    d = rows.setdefault(r["ID"], {"name": name})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1000.0 if u.startswith("n") else v * 1000.0 if u.startswith("m") else v
    if m.startswith("dram__bytes"):
        v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    d[m] = v
print(f"{'#':>3} {'kernel':34s} {'grid':>5} {'us':>7} {'tensor%':>8} {'sm%':>6} {'rd MB':>8} {'wr MB':>8} {'GB/s':>7} {'issue%':>7} {'regs':>5}")
tot = OrderedDict()
for i, d in enumerate(rows.values()):
    us = d["gpu__time_duration.sum"]
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    print(f"{i:3d} {d['name']:34s} {int(d.get('launch__grid_size', 0)):5d} {us:7.1f} "
          f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):8.1f} "
          f"{d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):6.1f} {rd:8.1f} {wr:8.1f} {(rd + wr) / us * 1e3:7.0f} "
          f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):7.1f} {int(d.get('launch__registers_per_thread', 0)):5d}")
    t = tot.setdefault(d["name"], [0.0, 0.0, 0.0, 0])
    t[0] += us; t[1] += rd; t[2] += wr; t[3] += 1
print()
all_us = sum(t[0] for t in tot.values())
for k, t in tot.items():
    print(f"  {k:34s} x{t[3]:<3d} {t[0]:8.1f} us {100 * t[0] / all_us:5.1f} %  {t[0] / tiles:6.2f} us/tile  read {t[1]:8.1f} MB  written {t[2]:8.1f} MB")
print(f"  total {all_us:.1f} us = {all_us / tiles:.2f} us per tile")
