"""One device-resident encode of the bench stream shape (44.1 kHz stereo multitone+noise) for ncu captures.
usage: python tools/prof_encode.py [seconds=592] [level=5] [reps=3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, flo_b200, synth_torch
SR, CH = 44100, 2
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 592
level = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = flo_b200.Context(0)
pcm = synth_torch.synth_pcm16_long(secs * SR, CH, SR, 0xF12, "multitone", 64, "cuda")
x = pcm.float() * (1 / 32768)
n = [x.numel()]
bound = ctx.output_bound(n, [SR], [CH])
out = torch.empty(bound, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    ctx.encode_batch_device([x.data_ptr()], n, [SR], [CH], [16], out.data_ptr(), bound, level=level)
t = ctx.last_timing()
print(f"{secs} s level {level}: encode kernel {t['encode_ms']:.3f} ms, device pass {t['device_ms']:.3f} ms")
