"""Timing of the analysis-metadata entries (row N4) on one hour of 44.1 kHz stereo f32 resident in HBM:
waveform peaks (50 per second) and EBU R128 integrated loudness (parity with the oracle: tests/test_gpu_analysis.py).  usage: python tools/bench_analysis.py [seconds=3600]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, flo_b200, synth_torch
from flo_b200 import analysis as fa
SR, CH = 44100, 2
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
pcm = synth_torch.synth_pcm16_long(secs * SR, CH, SR, 0xF12, "multitone", 64, "cuda")
x = pcm.float() * (1 / 32768)
n = x.numel()
cap = fa.peaks_count(n, SR, CH, 50)
peaks = torch.empty(cap, dtype=torch.float32, device="cuda")
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)
t_peaks = timed(lambda: fa.extract_waveform_peaks_device(x.data_ptr(), n, CH, SR, 50, peaks.data_ptr(), cap))
lufs = [None]
def loud(): lufs[0] = fa.integrated_loudness_device(x.data_ptr(), n, CH, SR)
t_loud = timed(loud)
out = {"audio_seconds": secs, "input_GB": n * 4 / 1e9, "peaks_ms": t_peaks, "peaks_GBps": n * 4 / t_peaks / 1e6, "n_peaks": cap,
       "loudness_ms": t_loud, "loudness_GBps_per_pass": n * 4 / t_loud / 1e6, "integrated_lufs": lufs[0]}
print(json.dumps(out))
