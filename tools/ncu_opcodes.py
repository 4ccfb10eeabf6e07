"""Instruction mix (executed warp instructions by SASS opcode) from an ncu source-page CSV dump."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
src_i = h.index("Source"); sass_i = h.index("Source", src_i + 1)
inst_i = h.index("Instructions Executed"); addr_i = h.index("Address")
seen = set(); mix = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) <= inst_i or not r[addr_i] or r[addr_i] in seen: continue
    seen.add(r[addr_i])
    try: n = int(r[inst_i] or 0)
    except ValueError: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[sass_i])
    op = m.group(2) if m else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "STS", "LDG", "STG", "SHF", "I2F", "F2I", "DADD", "ISETP")) else op.split(".")[0]
    mix[op] += n; tot += n
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0
print("total warp instructions", tot)
for op, n in mix.most_common(45):
    extra = f"  {n*scale:7.2f}/sample" if scale else ""
    print(f"{op:16s} {n:12d} {100*n/tot:5.1f}%{extra}")
