"""Host cost of a 4096-track host call (config-5 shape, pinned inputs): the Python mirror next to the bare C call.
usage: python tools/probe_many_tracks.py"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch, flo_b200
from flo_b200 import _lib
SR, CH, NTR, SECS = 8000, 1, 4096, 8
big = torch.empty(NTR*SECS*SR, dtype=torch.float32, pin_memory=True)
big.copy_((torch.rand(NTR*SECS*SR)-0.5)*0.1)
xs=[big[i*SECS*SR:(i+1)*SECS*SR] for i in range(NTR)]
ctx=flo_b200.Context(0)
specs=[flo_b200.TrackSpec(x.numpy(), SR, CH, 16, b"") for x in xs]
for _ in range(2):
    with ctx.encode_batch(specs,5,views=True) as r: pass
t0=time.perf_counter()
for _ in range(3):
    with ctx.encode_batch(specs,5,views=True) as r: nb=r.total_bytes()
print("python wrapper per call ms", (time.perf_counter()-t0)/3*1e3, ctx.last_timing())
arr=(_lib.Track*NTR)()
for i,x in enumerate(xs):
    arr[i].samples=x.data_ptr(); arr[i].n_interleaved=x.numel(); arr[i].sample_rate=SR; arr[i].channels=CH; arr[i].bit_depth=16
outs=(_lib.Out*NTR)()
def raw():
    _lib.check(ctx._L.flo_encode_batch(ctx._h, arr, NTR, 0, 5, outs))
    for i in range(NTR): ctx._L.flo_free(outs[i].data)
raw()
t0=time.perf_counter()
for _ in range(3): raw()
print("C call + frees per call ms", (time.perf_counter()-t0)/3*1e3, ctx.last_timing())
