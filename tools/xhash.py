"""Hashes of the encoder output for a few shapes (to compare experimental builds against the validated library).
usage: FLO_B200_SO=... python tools/xhash.py [levels=0,2,5]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, flo_b200, synth_torch
levels = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 2, 5]
ctx = flo_b200.Context(0)
cases = [("44k1 stereo 40s", [40 * 44100 + 1234], 44100, 2, "multitone", 64),
         ("44k1 mono 30s", [30 * 44100 + 7], 44100, 1, "multitone", 64),
         ("8k mono 64x3s", [3 * 8000 + 5] * 64, 8000, 1, "speech", 16),
         ("96k stereo 6s", [6 * 96000 + 99], 96000, 2, "sweep", 32),
         ("48k stereo 5s", [5 * 48000], 48000, 2, "multitone", 64)]
for name, tracks, sr, ch, kind, noise in cases:
    pcm = [synth_torch.synth_pcm16_long(n, ch, sr, 0xF30 + i, kind, noise, "cuda") for i, n in enumerate(tracks)]
    x = [p.float() * (1 / 32768) for p in pcm]
    n = [t.numel() for t in x]
    bound = ctx.output_bound(n, [sr] * len(n), [ch] * len(n))
    out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    for lv in levels:
        off, ln = ctx.encode_batch_device([t.data_ptr() for t in x], n, [sr] * len(n), [ch] * len(n), [16] * len(n), out.data_ptr(), bound, level=lv)
        h = hashlib.sha256()
        host = out.cpu().numpy()
        for o, l in zip(off, ln):
            h.update(host[int(o):int(o) + int(l)].tobytes())
        print(f"{name} L{lv} {int(sum(ln))} {h.hexdigest()[:16]}")
