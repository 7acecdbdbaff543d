"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares, not absolutes)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[h + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    v = v / 1e6 if u == "ns" else (v / 1e3 if u in ("us", "usecond") else v)
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^(bseg::|<unnamed>::)", "", name)
    d[name[:70]].append(v)
tot = sum(sum(v) for v in d.values())
print(f"total {tot:.1f} ms over {sum(len(v) for v in d.values())} launches")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} launches={len(v):5d} total_ms={sum(v):9.2f} share={sum(v) / tot:6.3f} avg_ms={sum(v) / len(v):8.4f}")
