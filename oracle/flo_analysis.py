"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's waveform-peak extraction
(libflo/src/core/analysis.rs:38-119, called by libflo::encode() through add_analysis_data_if_missing,
libflo/src/lib.rs:219-241).  Only tests/ may import this file; the product computes peaks in CUDA.

PARITY UNPINNED: the reference's tests for this function (libflo/tests/rust/analysis_tests.rs:3-67) assert
properties only (peaks in [0, 1], determinism, empty input) and hold no golden values, and there is no Rust
toolchain here to run the reference.  tests/test_oracle_golden.py checks those properties on the reference's own
test inputs and a few answers derived by hand from the source; every operation below is a single IEEE f32
operation (numpy float32), in the order the source performs them.
"""
import math

import numpy as np

F = np.float32


def _usize(v: float) -> int:
    """Rust `f64 as usize`: truncation toward zero, saturating, NaN -> 0"""
    if v != v:
        return 0
    if v >= 18446744073709551615.0:
        return (1 << 64) - 1
    return int(v) if v > 0 else 0


def extract_waveform_peaks(samples, channels: int, sample_rate: int, peaks_per_second: int) -> np.ndarray:
    x = np.asarray(samples, dtype=np.float32).reshape(-1)
    n = x.size
    if n == 0:                                                        # analysis.rs:44-50
        return np.zeros(0, np.float32)
    if channels == 0 or sample_rate == 0:
        raise OverflowError("capacity overflow")                      # Vec::with_capacity(usize::MAX), analysis.rs:54-56
    spp = float(sample_rate) / float(peaks_per_second) if peaks_per_second else math.inf      # :52
    total = _usize(math.ceil(n / (spp * channels)))                   # :53
    peaks = []
    for idx in range(total):
        start = _usize(idx * spp) * channels                          # :59, :62
        end = min(_usize((idx + 1.0) * spp) * channels, n)            # :60, :63
        if start >= n:
            break                                                     # :65-67
        w = x[start:end]
        if channels == 1:                                             # :72-78  fold(0.0, f32::max) over |s|; max skips NaN
            peak = np.fmax.reduce(np.abs(w), initial=F(0))
        elif channels == 2:                                           # :80-91  chunks_exact(2)
            pr = w[: (w.size // 2) * 2].reshape(-1, 2)
            l = np.fmax.reduce(np.abs(pr[:, 0]), initial=F(0))
            r = np.fmax.reduce(np.abs(pr[:, 1]), initial=F(0))
            peak = F(F(l + r) / F(2.0))
        else:                                                         # :93-99  chunks(channels): mean of each frame, no abs
            peak = F(0)
            for a in range(0, w.size, channels):
                ch = w[a:a + channels]
                s = F(0)
                for v in ch:                                          # sequential f32 sum
                    s = F(s + v)
                peak = np.fmax(peak, F(s / F(ch.size)))
        peaks.append(F(peak))
    out = np.array(peaks, dtype=np.float32)
    mx = np.fmax.reduce(out, initial=F(0)) if out.size else F(0)      # :104
    if mx > 0:
        with np.errstate(invalid="ignore"):
            out = (out / F(mx)).astype(np.float32)                    # :105-109
    return out
