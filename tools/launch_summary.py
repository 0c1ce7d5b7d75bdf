"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total time / share, for the
LAST `n_last` launches (= one step when the command runs identical steps; pass the per-step launch count).
    python tools/launch_summary.py launches.csv [n_last]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
ev = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    ev.append((re.sub(r"\(.*", "", r[ki])[:64], v))
if len(sys.argv) > 2:
    ev = ev[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for k, v in ev:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("launches %d, total %.1f us" % (len(ev), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s %4d %10.1f us %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
