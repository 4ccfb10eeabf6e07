"""Time the GPU lossless decoder on the bench workload (1 h of 44.1 kHz stereo, level 5): device-resident
decode (kernel times from CUDA events inside the library) and the host entry (H2D + decode + D2H)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import flo_b200 as fb
sys.path.insert(0, "tools")
import synth_torch

sr, ch = 44100, 2
seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
level = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = fb.default_context()
d_in = synth_torch.synth_pcm16_long(sr * seconds, ch, sr, 0xF10, "multitone", 64, "cuda")
n = d_in.numel()
bound = ctx.output_bound([n], [sr], [ch])
d_file = torch.empty(bound, dtype=torch.uint8, device="cuda")
offs, lens = ctx.encode_batch_device([d_in.data_ptr()], [n], [sr], [ch], [16], d_file.data_ptr(), bound, level=level, fmt=fb.FMT_PCM16)
flen = int(lens[0])
d_out = torch.empty(n, dtype=torch.float32, device="cuda")
res = {"seconds": seconds, "level": level, "file_bytes": flen, "pcm_bytes": n * 2}
ts = []
for it in range(6):
    cnt, info = ctx.decode_device(d_file.data_ptr() + int(offs[0]), flen, d_out.data_ptr(), n)
    ts.append(ctx.last_timing())
assert cnt == n
q = torch.trunc(torch.clamp(d_in.float() * (1.0 / 32768.0) * 32767.0, -32767, 32767))
res["round_trip_exact"] = bool(torch.equal(d_out, q * torch.tensor(1.0 / 32767.0, dtype=torch.float32, device="cuda")))
res["units_kernel_ms"] = sorted(t["encode_ms"] for t in ts[1:])[len(ts[1:]) // 2]
res["parse_ms"] = sorted(t["misc_ms"] for t in ts[1:])[len(ts[1:]) // 2]
res["device_pass_ms"] = sorted(t["device_ms"] for t in ts[1:])[len(ts[1:]) // 2]
res["decode_gbps_pcm"] = n * 2 / (res["units_kernel_ms"] * 1e-3) / 1e9
res["x_realtime"] = seconds / (res["device_pass_ms"] * 1e-3)
file_host = d_file[int(offs[0]):int(offs[0]) + flen].cpu().numpy()
he = []
for it in range(4):
    t0 = time.perf_counter()
    out, info = ctx.decode(file_host)
    he.append((time.perf_counter() - t0) * 1e3)
    lt = ctx.last_timing()
    del out
res["host_entry_ms"] = min(he[1:])
res["host_h2d_ms"] = lt["h2d_ms"]; res["host_d2h_ms"] = lt["d2h_ms"]
pinned = torch.from_numpy(file_host).pin_memory().numpy()          # same call with the file bytes in page-locked memory
he = []
for it in range(4):
    t0 = time.perf_counter()
    out, info = ctx.decode(pinned)
    he.append((time.perf_counter() - t0) * 1e3)
    lt = ctx.last_timing()
    del out
res["host_entry_pinned_ms"] = min(he[1:])
he = []
for it in range(4):                                                  # flo_decode_i16: half the bytes come back
    t0 = time.perf_counter()
    out, info = ctx.decode_i16(pinned)
    he.append((time.perf_counter() - t0) * 1e3)
    lt16 = ctx.last_timing()
    del out
res["host_entry_pinned_i16_ms"] = min(he[1:]); res["i16_units_kernel_ms"] = lt16["encode_ms"]; res["i16_d2h_ms"] = lt16["d2h_ms"]
res["host_pinned_h2d_ms"] = lt["h2d_ms"]
print(json.dumps(res))
