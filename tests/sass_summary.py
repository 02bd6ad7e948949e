"""Per-kernel SASS evidence for profiles/ (not a pytest module): disassembles libflairb200.so with cuobjdump and counts, per
kernel, the Blackwell instructions that prove what each one is built on -- UTCHMMA (tcgen05.mma; .2CTA = cta_group::2),
LDTM (tcgen05.ld), UTMALDG (TMA tensor load), UBLKCP (1-D bulk copy), LDGSTS (cp.async), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), REDG/ATOMG (global atomics), MATCH (warp match). Runs without a GPU.

    python tests/sass_summary.py > profiles/r02_sass_summary.txt
"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "flair-1_b200" / "libflairb200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = iter(demangle)
PATTERNS = OrderedDict([("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"),
                        ("UBLKCP", r"\bUBLKCP"), ("LDGSTS", r"\bLDGSTS"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"),
                        ("ATOM/RED", r"\b(ATOMG|REDG|ATOMS|RED)\b"), ("MATCH", r"\bMATCH")])
print(f"{lib.name}: cuobjdump -sass, instruction counts per kernel (static occurrences in the code, sm_100a)")
print(f"{'kernel':64s} " + " ".join(f"{k:>12s}" for k in PATTERNS))
tot = {k: 0 for k in PATTERNS}
for block in out.split("Function : ")[1:]:
    name = next(names)
    name = name.replace("fb::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r">\(.*", ">", name) if "<" in name else re.sub(r"\(.*", "", name)
    name = name.replace("fb::", "")
    counts = {k: len(re.findall(p, block)) for k, p in PATTERNS.items()}
    for k, v in counts.items():
        tot[k] += v
    print(f"{name[:64]:64s} " + " ".join(f"{counts[k]:12d}" for k in PATTERNS))
print(f"{'total':64s} " + " ".join(f"{tot[k]:12d}" for k in PATTERNS))
