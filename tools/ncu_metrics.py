"""Print metrics of an .ncu-rep whose names match any of the given regexes: tools/ncu_metrics.py rep.ncu-rep re1 re2 ..."""
import csv, re, subprocess, sys
rep, pats = sys.argv[1], [re.compile(p) for p in sys.argv[2:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, r):
        if any(p.search(h) for p in pats):
            print(f"  {h} [{u}] = {v}")
