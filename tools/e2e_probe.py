"""Where an end-to-end host call spends its time: wall time of flo_encode_batch (pinned host f32 in, pinned .flo out)
next to the library's own event timings.  usage: python tools/e2e_probe.py [seconds=3600] [pageable]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, flo_b200, synth_torch
SR, CH = 44100, 2
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
pcm = synth_torch.synth_pcm16_long(secs * SR, CH, SR, 0xF12, "multitone", 64, "cuda")
x = (pcm.float() * (1 / 32768))
pageable = len(sys.argv) > 2 and sys.argv[2] == "pageable"
h = torch.empty(x.numel(), dtype=torch.float32, pin_memory=not pageable); h.copy_(x); torch.cuda.synchronize()
ctx = flo_b200.Context(0)
specs = [flo_b200.TrackSpec(h.numpy(), SR, CH, 16, b"")]
for i in range(5):
    t0 = time.perf_counter()
    with ctx.encode_batch(specs, 5, views=True) as res:
        nb = res.total_bytes()
    dt = (time.perf_counter() - t0) * 1e3
    print(f"call {dt:.2f} ms, {nb} bytes;", {k: round(v, 3) for k, v in ctx.last_timing().items()})
