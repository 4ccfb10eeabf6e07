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
# where the bulk push spends its time: the C entry alone (one device pass + re-serialisation), then the Python mirror
import ctypes as C
from flo_b200 import _lib
ctx = flo_b200.default_context(0)
xc = np.ascontiguousarray(x)
ts = []
for _ in range(3):
    out, out_len, offs, nfr = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_uint32()
    t0 = time.perf_counter()
    _lib.check(ctx._L.flo_stream_encode_frames(ctx._h, xc.ctypes.data_as(C.c_void_p), xc.size, sr, ch, 16, 5, C.byref(out), C.byref(out_len), C.byref(offs), C.byref(nfr)))
    ts.append((time.perf_counter() - t0) * 1e3)
    ctx._L.flo_free(out); ctx._L.flo_free(offs)
print(json.dumps({"flo_stream_encode_frames_ms_600_frames": min(ts), "frames": nfr.value, "bytes": out_len.value}))
