"""Latency / throughput of flo_b200.StreamingEncoder (row N3): one 1-second 44.1 kHz stereo frame per push
(the live case) and 600 frames in one push (one device pass)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import flo_b200
from helpers import pcm16_to_f32, synth_pcm16

sr, ch = 44100, 2
x = pcm16_to_f32(synth_pcm16(sr * 600, ch, sr, seed=77))
enc = flo_b200.StreamingEncoder(sr, ch, 16)
enc.push_samples(x[: sr * ch]); enc.next_frame()                       # warm-up
lat = []
for i in range(1, 41):
    t0 = time.perf_counter()
    enc.push_samples(x[i * sr * ch:(i + 1) * sr * ch])
    fr = enc.next_frame()
    lat.append((time.perf_counter() - t0) * 1e3)
assert fr is not None and fr.samples == sr
enc2 = flo_b200.StreamingEncoder(sr, ch, 16)
enc2.push_samples(x[: 4 * sr * ch]); enc2.finalize()
enc2 = flo_b200.StreamingEncoder(sr, ch, 16)
t0 = time.perf_counter()
enc2.push_samples(x)
dt = time.perf_counter() - t0
n = enc2.pending_frames()
out = enc2.finalize(b"")
print(json.dumps({"per_frame_push_ms_median": float(np.median(lat)), "per_frame_push_ms_min": float(min(lat)),
                  "bulk_frames": n, "bulk_push_ms": dt * 1e3, "bulk_x_realtime": n / dt, "file_bytes": len(out)}))
