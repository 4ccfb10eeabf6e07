"""Small mixed workload for compute-sanitizer (memcheck / racecheck): every kernel variant, both plane
placements, levels 2/5/9, mono / stereo / 3 channels, ragged tails; results are checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
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
print("sanitize_case ok")
