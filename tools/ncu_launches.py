"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns": v /= 1e3
        elif r[ui] == "ms": v *= 1e3
        d.setdefault(r[ki].split("(")[0][:70], []).append(v)
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print(f"{k:70s} n={len(v):3d} mean={sum(v)/len(v):10.1f} us  share={100*sum(v)/tot:5.1f}%")
