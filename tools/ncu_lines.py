"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: python tools/ncu_lines.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
# first "Source" col is the CUDA line text, second is SASS
src_i = h.index("Source"); sass_i = h.index("Source", src_i + 1)
samp_i = col["# Samples"]; inst_i = col["Instructions Executed"]
stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) <= inst_i: continue
    try:
        ln = r[0]; s = int(r[samp_i] or 0); n = int(r[inst_i] or 0)
    except ValueError:
        continue
    a = agg[ln]; a[0] += s; a[1] += n; a[2] = r[src_i][:110]
    for nme, i in stall_cols:
        try: a[3][nme] += int(r[i] or 0)
        except ValueError: pass
    tot_s += s; tot_i += n
print(f"total samples {tot_s}, warp instructions {tot_i}")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k[6:]}:{v}" for k, v in a[3].most_common(3))
    print(f"L{ln:>5} samp {100*a[0]/max(tot_s,1):5.1f}%  inst {100*a[1]/max(tot_i,1):5.1f}%  [{st}]  {a[2].strip()}")
