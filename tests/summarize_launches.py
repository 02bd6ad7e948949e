"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of tests/prof_forward.py.

    python tests/summarize_launches.py gpurun_out/launches.csv [launches_per_pass] [tiles]

Prints the last pass launch by launch (us, us per tile) and the per-family totals. Not a pytest module.
"""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit.startswith("n") else v * 1000.0 if unit.startswith("m") else v
        name = re.sub(r"^void\s+", "", r["Kernel Name"])
        name = re.sub(r"\(.*", "", name).replace("unnamed>::", "").replace("fb::", "")
        rows.append((name, r["Grid Size"], us))
    return rows


def main():
    path = sys.argv[1]
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 49
    tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 37
    rows = load(path)
    last = rows[-per:]
    fam = OrderedDict()
    tot = 0.0
    for i, (name, grid, us) in enumerate(last):
        tot += us
        fam[name] = fam.get(name, 0.0) + us
        print(f"{i:3d} {name:42s} {grid:14s} {us:8.1f} us {us / tiles:6.2f} us/tile")
    print(f"total {tot:.1f} us = {tot / tiles:.2f} us per tile ({len(rows)} launches in file, {per} per pass)")
    for k, v in fam.items():
        print(f"  {k:42s} {v:8.1f} us {100 * v / tot:5.1f} %")


if __name__ == "__main__":
    main()
