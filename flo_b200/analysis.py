"""Host-side mirror of `libflo_audio::core::analysis::extract_waveform_peaks` (libflo/src/core/analysis.rs:38-119),
the first piece of the analysis metadata `libflo::encode()` attaches (libflo/src/lib.rs:97-117, 219-241; SURVEY 8f
row N4).  The work happens in the CUDA library (flo_waveform_peaks); nothing here computes peaks on the CPU."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .encoder import default_context


@dataclass
class WaveformData:
    """core/metadata.rs WaveformData: peaks_per_second, peaks (normalised 0..1), channels"""
    peaks_per_second: int
    peaks: np.ndarray
    channels: int


def extract_waveform_peaks(samples, channels: int, sample_rate: int, peaks_per_second: int = 50, ctx=None) -> WaveformData:
    """Same arguments and order as the reference (samples, channels, sample_rate, peaks_per_second); 50 peaks per
    second is what libflo::encode() asks for (lib.rs:110)."""
    x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
    ctx = ctx or default_context(0)
    out, n = C.c_void_p(), C.c_size_t()
    _lib.check(ctx._L.flo_waveform_peaks(ctx._h, x.ctypes.data_as(C.c_void_p), x.size, int(sample_rate), int(channels), int(peaks_per_second),
                                        C.byref(out), C.byref(n)))
    try:
        peaks = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
    finally:
        if out.value:
            ctx._L.flo_free(out)
    return WaveformData(int(peaks_per_second), peaks, int(channels))


def extract_waveform_peaks_device(d_samples: int, n_interleaved: int, channels: int, sample_rate: int, peaks_per_second: int,
                                  d_peaks: int, capacity: int, ctx=None) -> int:
    """Device pointers in, device peaks out; returns the number of peaks written."""
    ctx = ctx or default_context(0)
    n = C.c_size_t()
    _lib.check(ctx._L.flo_waveform_peaks_device(ctx._h, C.c_void_p(d_samples), int(n_interleaved), int(sample_rate), int(channels),
                                               int(peaks_per_second), C.c_void_p(d_peaks), int(capacity), C.byref(n)))
    return int(n.value)


def peaks_count(n_interleaved: int, sample_rate: int, channels: int, peaks_per_second: int) -> int:
    return int(_lib.lib().flo_waveform_peaks_count(int(n_interleaved), int(sample_rate), int(channels), int(peaks_per_second)))


@dataclass
class LoudnessMetrics:
    """core/ebu_r128.rs LoudnessMetrics; only integrated_lufs is computed (the one value libflo::encode() stores)"""
    integrated_lufs: float
    loudness_range_lu: float = None
    true_peak_dbtp: float = None
    sample_peak_dbfs: float = None


def compute_ebu_r128_loudness(samples, channels: int, sample_rate: int, ctx=None) -> LoudnessMetrics:
    """Same arguments as the reference (libflo/src/core/ebu_r128.rs:182-186)."""
    x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
    ctx = ctx or default_context(0)
    v = C.c_double()
    _lib.check(ctx._L.flo_integrated_loudness(ctx._h, x.ctypes.data_as(C.c_void_p), x.size, int(sample_rate), int(channels), C.byref(v)))
    return LoudnessMetrics(float(v.value))


def integrated_loudness_device(d_samples: int, n_interleaved: int, channels: int, sample_rate: int, ctx=None) -> float:
    ctx = ctx or default_context(0)
    v = C.c_double()
    _lib.check(ctx._L.flo_integrated_loudness_device(ctx._h, C.c_void_p(d_samples), int(n_interleaved), int(sample_rate), int(channels), C.byref(v)))
    return float(v.value)
