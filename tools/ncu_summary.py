"""Summaries of an ncu --set full capture for profiles/: key metrics (raw page) and per-source-line hotspots
(source page, CUDA-line rows only).  usage: python tools/ncu_summary.py rep.ncu-rep out_prefix "description" """
import collections, csv, io, subprocess, sys
rep, prefix, desc = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
units = rows[1] if len(rows) > 2 else [""] * len(h)
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
with open(prefix + "_key_metrics.txt", "w") as f:
    f.write(desc + "\n")
    col = {n: i for i, n in enumerate(h)}
    for n in want:
        if n in col:
            f.write(f"{n:75s} {v[col[n]]:>18s} {units[col[n]]}\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True, errors="replace").stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
src_i = h.index("Source")
samp_i, inst_i = col["# Samples"], col["Instructions Executed"]
stall = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
lines, tot_s, tot_i, tot_st = [], 0, 0, collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= inst_i or r[0] == "":
        continue                                  # SASS rows repeat their CUDA line's totals
    try:
        s, n = int(r[samp_i] or 0), int(r[inst_i] or 0)
    except ValueError:
        continue
    c = collections.Counter({k[6:]: int(r[i] or 0) for k, i in stall})
    lines.append((s, n, r[0], r[src_i].strip()[:110], c))
    tot_s += s; tot_i += n; tot_st.update(c)
with open(prefix + "_source_hotspots.txt", "w") as f:
    f.write(desc + "\n")
    f.write(f"total warp samples {tot_s}, warp instructions {tot_i}\n")
    f.write("stall reasons, share of samples: " + ", ".join(f"{k} {100 * v / max(tot_s, 1):.1f}%" for k, v in tot_st.most_common(10)) + "\n")
    for s, n, ln, text, c in sorted(lines, key=lambda x: -x[0])[:45]:
        top = ", ".join(f"{k}:{v}" for k, v in c.most_common(3))
        f.write(f"L{ln:>5} samp {100 * s / max(tot_s, 1):5.1f}%  inst {100 * n / max(tot_i, 1):5.1f}%  [{top}]  {text}\n")
print(open(prefix + "_key_metrics.txt").read())
