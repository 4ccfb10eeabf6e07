"""Attribute executed SASS opcodes to CUDA source lines from an ncu source-page CSV (cuda,sass view).
usage: python tools/ncu_ops_by_line.py dump.csv scale OP [OP ...]   (scale = 32 / samples for per-sample counts)"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
scale = float(sys.argv[2]); ops = sys.argv[3:]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
src_i = h.index("Source"); sass_i = h.index("Source", src_i + 1); inst_i = h.index("Instructions Executed")
byop = collections.defaultdict(collections.Counter)
cur = "?"
for r in rows[hi + 1:]:
    if len(r) <= inst_i: continue
    if r[0]:
        cur = r[0] + ": " + r[src_i].strip()[:100]
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[sass_i])
    if not m: continue
    try: n = int(r[inst_i] or 0)
    except ValueError: continue
    byop[m.group(2).split(".")[0]][cur] += n
for op in ops:
    tot = sum(byop[op].values())
    print(f"== {op} total {tot*scale:.2f}/sample")
    for k, v in byop[op].most_common(8): print(f"   {v*scale:6.2f}  {k}")
