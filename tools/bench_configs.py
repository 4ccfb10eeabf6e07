"""Device-resident timings of the BASELINE configs other than the headline one (parity-test cases, not bench lines):
config 3 shard (20 x 180 s 44.1k stereo), config 4 (96 kHz stereo, level 9), config 5 (8 kHz mono short tracks)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, flo_b200, synth_torch

ctx = flo_b200.Context(0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def run(name, tracks, sr, ch, level, kind, noise, bits=16):
    pcm = [synth_torch.synth_pcm16_long(n, ch, sr, 0xF20 + i, kind, noise, "cuda") for i, n in enumerate(tracks)]
    x = [p.float() * (1 / 32768) for p in pcm]
    n = [t.numel() for t in x]
    bound = ctx.output_bound(n, [sr] * len(n), [ch] * len(n))
    out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    best = None
    for _ in range(4):
        off, ln = ctx.encode_batch_device([t.data_ptr() for t in x], n, [sr] * len(n), [ch] * len(n), [bits] * len(n),
                                          out.data_ptr(), bound, level=level)
        t = ctx.last_timing()
        best = t if best is None or t["device_ms"] < best["device_ms"] else best
    tot = sum(n)
    secs = sum(tracks) / sr
    alg = 4.0 * tot + float(ln.sum())
    print(json.dumps({"config": name, "tracks": len(tracks), "audio_s": secs, "level": level,
                      "device_ms": round(best["device_ms"], 3), "encode_ms": round(best["encode_ms"], 3),
                      "pcm_GBps": round(2.0 * tot / best["device_ms"] / 1e6, 2), "x_realtime": round(secs / best["device_ms"] * 1e3),
                      "pct_hbm_peak_encode_kernel": round(100 * alg / best["encode_ms"] / 1e6 / peak, 2),
                      "ratio": round(2.0 * tot / float(ln.sum()), 3)}))


run("C3 shard: 20 x 180 s 44.1k stereo L5", [180 * 44100] * 20, 44100, 2, 5, "multitone", 64)
run("C4: 600 s 96k stereo (24-bit header) L9", [600 * 96000], 96000, 2, 9, "sweep", 32, bits=24)
run("C5: 4096 x 8 s 8k mono L5", [8 * 8000] * 4096, 8000, 1, 5, "speech", 16)
run("48k stereo 600 s L5", [600 * 48000], 48000, 2, 5, "multitone", 64)
