"""Host time of one device-entry call next to its device span, for 1 track and for 20 tracks of the same total length.
usage: python tools/call_overhead.py"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch, flo_b200, synth_torch
from flo_b200 import _lib
SR, CH = 44100, 2
ctx = flo_b200.Context(0)
for ntr in (1, 20):
    secs = 3600 // ntr
    xs = [synth_torch.synth_pcm16_long(secs * SR, CH, SR, 0xF12 + t, "multitone", 64, "cuda").float() * (1 / 32768) for t in range(ntr)]
    n = [x.numel() for x in xs]
    bound = ctx.output_bound(n, [SR] * ntr, [CH] * ntr)
    out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    ptrs = [x.data_ptr() for x in xs]
    arr = (_lib.Track * ntr)()
    for i in range(ntr):
        arr[i].samples = ptrs[i]; arr[i].n_interleaved = n[i]; arr[i].sample_rate = SR; arr[i].channels = CH; arr[i].bit_depth = 16
    off = np.zeros(ntr, np.uint64); ln = np.zeros(ntr, np.uint64)
    def raw():
        _lib.check(ctx._L.flo_encode_batch_device(ctx._h, arr, ntr, 0, 5, C.c_void_p(out.data_ptr()), bound,
                                                 off.ctypes.data_as(C.POINTER(C.c_uint64)), ln.ctypes.data_as(C.POINTER(C.c_uint64))))
    def wrapped():
        ctx.encode_batch_device(ptrs, n, [SR] * ntr, [CH] * ntr, [16] * ntr, out.data_ptr(), bound, level=5)
    for name, f in (("C call", raw), ("Python wrapper", wrapped)):
        for _ in range(3): f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20 * 1e3
        print(f"{ntr:2d} tracks, {name}: {dt:.3f} ms per call, device span {ctx.last_timing()['device_ms']:.3f} ms, encode kernel {ctx.last_timing()['encode_ms']:.3f} ms")
t0 = time.perf_counter()
for _ in range(200): ctx.last_timing()
print(f"last_timing(): {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per query")
t0 = time.perf_counter()
for _ in range(200): [int(v) for v in ln]
print(f"lengths to a list: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
