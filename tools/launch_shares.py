"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[mu], 1.0)
    name = r[kn].split("(")[0]
    tot[name] += v; cnt[name] += 1
ours = {k: v for k, v in tot.items() if "uvgb::" in k}
s = sum(ours.values())
print("kernel | launches | total_us | share_of_our_kernels")
for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
    print("%s | %d | %.1f | %.1f%%" % (k.replace("void ", ""), cnt[k], v, 100 * v / s))
other = sum(v for k, v in tot.items() if "uvgb::" not in k)
print("(other kernels in the process - torch fills / copies / flush: %.1f us)" % other)
