"""Table of kernel durations from tests/halo_skip_probe.py run under an ncu launch list (not a pytest module).

    python tests/summarize_skip_probe.py gpurun_out/skip_launches.csv gpurun_out/skip_order.log
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from summarize_launches import load  # noqa: E402

rows = load(sys.argv[1])
order = [ln.strip().split(" | ") for ln in open(sys.argv[2]) if ln.startswith("LAUNCH")]
assert len(rows) == len(order), (len(rows), len(order))
res, shapes, masks = {}, [], []
for (name, grid, us), (_, shape, mask) in zip(rows, order):
    m = int(mask.split()[1])
    res[(shape, m)] = us  # the second launch of a pair overwrites the first
    if shape not in shapes:
        shapes.append(shape)
    if m not in masks:
        masks.append(m)
print("mask bits: 1 = no producer copies, 2 = no MMAs, 4 = no epilogue stores, 8 = no epilogue; us per launch")
print("shape".ljust(34) + "".join(f"m{m:<7d}" for m in masks))
for s in shapes:
    print(s.ljust(34) + "".join(f"{res[(s, m)]:<8.1f}" for m in masks))
