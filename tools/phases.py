"""Per-phase SM-clock breakdown of the frame-encode kernel (ingest / analysis / look-back / pack) per level."""
import sys, json
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tools'))
import torch, flo_b200, synth_torch
SR,CH=44100,2
ctx=flo_b200.Context(0)
secs=int(sys.argv[1]) if len(sys.argv)>1 else 1184
pcm=synth_torch.synth_pcm16_long(secs*SR,CH,SR,0xF12,"multitone",64,"cuda")
x=pcm.float()*(1/32768)
n=[x.numel()]; bound=ctx.output_bound(n,[SR],[CH]); out=torch.empty(bound,dtype=torch.uint8,device="cuda")
levels=[int(v) for v in sys.argv[2].split(',')] if len(sys.argv)>2 else (0,2,5,9)
for lv in levels:
    for _ in range(3): ctx.encode_batch_device([x.data_ptr()],n,[SR],[CH],[16],out.data_ptr(),bound,level=lv)
    t=ctx.last_timing(); c=ctx.last_counters()["phase_clocks"]; f=secs
    print(f"level {lv}: kernel {t['encode_ms']:.3f} ms; per-frame clocks: " + ", ".join(f"{k} {v/f:.0f}" for k,v in c.items()))
t = ctx.last_timing()
print("last call timing (ms):", {k: round(v, 3) if isinstance(v, float) else v for k, v in t.items()})
