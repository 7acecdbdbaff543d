"""Per-kernel SASS opcode histogram of libbseg.so (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma, LDTM / STTM
= tcgen05.ld / st, UTMALDG = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier, MUFU.EX2 ...).
usage: python tools/sass_histogram.py [path/to/libbseg.so] > profiles/rNN_sass_opcode_histogram.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "beach_seg_b200" / "libbseg.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "SYNCS", "UCGABAR", "MUFU", "HMMA", "FADD2",
       "FFMA2", "F2FP", "USETMAXREG", "ELECT", "REDG", "ATOMG", "ATOMS", "RED.", "LDGSTS")
kern, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        per[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        per[kern][m.group(1)] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n


print(f"# SASS opcode histogram of {Path(so).name} (cuobjdump -sass), instructions per kernel; only the opcodes that identify")
print("# the Blackwell paths (tcgen05 / TMEM / TMA / mbarrier), the special-function and packed-fp32 pipes and atomics")
tot = collections.Counter()
for k, c in per.items():
    n = sum(c.values())
    keys = {op: v for op, v in c.items() if any(op.startswith(p) for p in KEY)}
    agg = collections.Counter()
    for op, v in keys.items():
        base = op if op.startswith(("UTCHMMA", "UTMALDG", "MUFU", "SYNCS", "UTCBAR", "LDTM", "STTM")) else op.split(".")[0]
        agg[base] += v
        tot[base] += v
    if not agg:
        continue
    print(f"\n{demangle(k)}  [{n} instructions]")
    print("   " + "  ".join(f"{op}:{v}" for op, v in sorted(agg.items())))
print("\n# totals over the library")
print("   " + "  ".join(f"{op}:{v}" for op, v in sorted(tot.items())))
