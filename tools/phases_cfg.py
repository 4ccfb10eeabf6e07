"""Per-phase SM clocks for a BASELINE config shape (needs a -DFLO_PHASE_CLOCKS build).  usage: phases_cfg.py {3,4,5}"""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tools'))
import argparse, torch, flo_b200, synth_torch
sys.argv=[sys.argv[0]]+sys.argv[1:]
import importlib.util
spec=importlib.util.spec_from_file_location("bench", os.path.join(ROOT,"bench.py")); bench=importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
cid=int(sys.argv[1]) if len(sys.argv)>1 else 5
cc=bench.CONFIGS[cid]
lens=bench.track_lengths(cid, argparse.Namespace(seconds=600, tracks=20), 1)
dev=torch.device("cuda",0)
pcs=[synth_torch.synth_pcm16_long(n, cc["ch"], cc["sr"], 0xF12+131*i, cc["kind"], cc["noise"], dev) for i,n in enumerate(lens)]
xs=[p.to(torch.float32)*(1.0/32768.0) for p in pcs]
nn=[int(t.numel()) for t in xs]
ctx=flo_b200.Context(0)
bb=ctx.output_bound(nn,[cc["sr"]]*len(nn),[cc["ch"]]*len(nn)); oo=torch.empty(bb,dtype=torch.uint8,device=dev)
for _ in range(3):
    ctx.encode_batch_device([t.data_ptr() for t in xs],nn,[cc["sr"]]*len(nn),[cc["ch"]]*len(nn),[cc["bits"]]*len(nn),oo.data_ptr(),bb,level=cc["level"])
t=ctx.last_timing(); c=ctx.last_counters()["phase_clocks"]; f=sum(-(-n//cc["sr"]) for n in lens)
cn=ctx.last_counters(); print({k:v for k,v in cn.items() if k!="phase_clocks"})
print(f"config {cid}: kernel {t['encode_ms']:.3f} ms, {f} frames; per-frame clocks: "+", ".join(f"{k} {v/f:.0f}" for k,v in c.items()))
