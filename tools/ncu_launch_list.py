"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total time, share.
Usage: python tools/ncu_launch_list.py launches.csv [> profiles/x.txt]"""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= max(ik, im, iv) or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu].strip().replace("second", "s").replace("n", "n"), 1e-3)
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"\b(gk|at|cub|thrust)::", "", name)
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values())
print("%-112s %6s %12s %7s" % ("kernel", "count", "total us", "share"))
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-112s %6d %12.1f %6.1f%%" % (name[-112:], c, us, 100.0 * us / total))
print("total %.1f us over %d launches" % (total, sum(v[0] for v in agg.values())))
