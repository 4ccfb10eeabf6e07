"""(Test infrastructure, run by hand on a GPU box: `python tests/sanitize_case.py`, optionally under compute-sanitizer.)
Small mixed workload for compute-sanitizer (memcheck / racecheck): every kernel variant, both plane
placements, levels 2/5/9, mono / stereo / 3 channels, ragged tails; results are checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from helpers import oracle, pcm16_to_f32, synth_pcm16
import flo_b200

ctx = flo_b200.Context(0)
cases = [(2 * 8000 + 5, 1, 8000, 5), (44100 + 333, 2, 44100, 5), (22050 + 7, 2, 22050, 9), (3000, 3, 1000, 5),
         (44100 + 1, 1, 44100, 2), (2 * 96000 // 8, 2, 96000, 9)]
for variant in (None, "512", "256", "128"):
    if variant is None:
        os.environ.pop("FLO_B200_VARIANT", None)
    else:
        os.environ["FLO_B200_VARIANT"] = variant
    for n, ch, sr, lvl in cases:
        pcm = synth_pcm16(n, ch, sr, seed=n + ch)
        x = pcm16_to_f32(pcm)
        got = flo_b200.Encoder(sr, ch, 16, context=ctx).with_compression(lvl).encode(x, b"m")
        assert got == oracle.encode(x, sr, ch, 16, lvl, b"m"), (variant, n, ch, sr, lvl)
os.environ.pop("FLO_B200_VARIANT", None)
# decoder, analysis entries (round 2): decode round trip, waveform peaks, R128 loudness incl. ragged and many-channel inputs
from flo_b200 import analysis as fa
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import flo_analysis
for n, ch, sr, lvl in cases:
    x = pcm16_to_f32(synth_pcm16(n, ch, sr, seed=n + ch))
    img = flo_b200.Encoder(sr, ch, 16, context=ctx).with_compression(lvl).encode(x, b"")
    dec = flo_b200.Decoder().decode(img)
    assert np.array_equal(dec, oracle.decode(img))
    pk = fa.extract_waveform_peaks(x, ch, sr, 50, ctx=ctx).peaks
    assert np.array_equal(pk.view(np.uint32), flo_analysis.extract_waveform_peaks(x, ch, sr, 50).view(np.uint32))
    if sr >= 8000:                                                  # below ~3.4 kHz the reference's shelf filter is unstable
        lu = fa.compute_ebu_r128_loudness(x, ch, sr, ctx=ctx).integrated_lufs
        assert abs(lu - oracle.r128_integrated_lufs(x, ch, sr)) < 1e-9
x = pcm16_to_f32(synth_pcm16(9 * 5000 + 3, 9, 8000, seed=5))        # more than 8 channels: the unstaged K-weighting kernel
assert abs(fa.compute_ebu_r128_loudness(x, 9, 8000, ctx=ctx).integrated_lufs - oracle.r128_integrated_lufs(x, 9, 8000)) < 1e-9
print("sanitize_case ok")
