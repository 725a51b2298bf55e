"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, summed time, share.
    python scripts/summarize_launches.py profiles/r02_launches_bench_parity_config_2.csv > ..._summary.txt"""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv:
        continue
    t = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1e-3)
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name)
    a = agg[name]
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms summed (ncu serialises launches and runs them cold: shares are meaningful, absolutes are not)")
print("# launches | us total | share | kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[0]:5d} {v[1]:10.1f} {v[1] / tot:6.3f}  {k[:150]}")
