"""ctypes binding of libflo_b200.so (include/flo_b200.h).  No fallback: a missing library or a
missing GPU is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FLO_B200_SO selects an experimental build of the same library (see build.py); never a different implementation
SO_PATH = os.environ.get("FLO_B200_SO") or os.path.join(_HERE, "libflo_b200.so")

FMT_F32, FMT_PCM16, FMT_U8, FMT_S32 = 0, 1, 2, 3


class FloError(RuntimeError):
    """Err(String) of the reference's FloResult (libflo/src/core/types.rs:281)."""


class Track(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("n_interleaved", C.c_size_t), ("sample_rate", C.c_uint32),
                ("channels", C.c_uint8), ("bit_depth", C.c_uint8), ("meta", C.c_void_p), ("meta_len", C.c_size_t)]


class Out(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t)]


class Info(C.Structure):
    """flo_info: header fields of a decoded file."""
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint8), ("bit_depth", C.c_uint8), ("level", C.c_uint8),
                ("version_major", C.c_uint8), ("total_samples", C.c_uint64), ("decoded_frames", C.c_uint64),
                ("n_frames", C.c_uint32), ("data_crc32", C.c_uint32), ("meta_offset", C.c_uint64), ("meta_size", C.c_uint64)]


class CandReport(C.Structure):
    _fields_ = [("k", C.c_int32), ("pad", C.c_int32), ("size", C.c_int64)]


# every symbol include/flo_b200.h declares
EXPORTS = [
    "flo_ctx_create", "flo_ctx_destroy", "flo_encode", "flo_encode_pcm16", "flo_encode_batch",
    "flo_encode_batch_device", "flo_output_bound", "flo_ctx_set_stream", "flo_ctx_last_timing",
    "flo_ctx_last_counters", "flo_ctx_enable_report", "flo_ctx_read_report", "flo_host_alloc", "flo_host_free", "flo_free",
    "flo_last_error", "flo_version", "flo_device_count", "flo_decode", "flo_decode_device", "flo_stream_encode_frames",
    "flo_waveform_peaks", "flo_waveform_peaks_device", "flo_waveform_peaks_count",
    "flo_integrated_loudness", "flo_integrated_loudness_device", "flo_decode_i16", "flo_decode_i16_device",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise FloError(f"{SO_PATH} is missing: build it with `python -m flo_b200.build` "
                       "(flo_b200 has no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, sz, u8, u32, u64p = C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.POINTER(C.c_uint64)
    L.flo_ctx_create.restype = C.c_int
    L.flo_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.flo_ctx_destroy.restype = None
    L.flo_ctx_destroy.argtypes = [vp]
    enc = [vp, vp, sz, u32, u8, u8, u8, vp, sz, C.POINTER(vp), C.POINTER(sz)]
    L.flo_encode.restype = C.c_int
    L.flo_encode.argtypes = enc
    L.flo_encode_pcm16.restype = C.c_int
    L.flo_encode_pcm16.argtypes = enc
    L.flo_encode_batch.restype = C.c_int
    L.flo_encode_batch.argtypes = [vp, C.POINTER(Track), sz, C.c_int, u8, C.POINTER(Out)]
    L.flo_encode_batch_device.restype = C.c_int
    L.flo_encode_batch_device.argtypes = [vp, C.POINTER(Track), sz, C.c_int, u8, vp, sz, u64p, u64p]
    L.flo_output_bound.restype = sz
    L.flo_output_bound.argtypes = [C.POINTER(Track), sz]
    L.flo_ctx_set_stream.restype = C.c_int
    L.flo_ctx_set_stream.argtypes = [vp, vp]
    L.flo_ctx_last_timing.restype = C.c_int
    L.flo_ctx_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(u32)]
    L.flo_ctx_last_counters.restype = C.c_int
    L.flo_ctx_last_counters.argtypes = [vp, u64p]
    L.flo_ctx_enable_report.restype = C.c_int
    L.flo_ctx_enable_report.argtypes = [vp, C.c_int]
    L.flo_ctx_read_report.restype = C.c_int
    L.flo_ctx_read_report.argtypes = [vp, u32, u32, C.POINTER(CandReport)]
    L.flo_decode.restype = C.c_int
    L.flo_decode.argtypes = [vp, vp, sz, C.POINTER(vp), C.POINTER(sz), C.POINTER(Info)]
    L.flo_decode_device.restype = C.c_int
    L.flo_decode_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(Info)]
    L.flo_decode_i16.restype = C.c_int
    L.flo_decode_i16.argtypes = [vp, vp, sz, C.POINTER(vp), C.POINTER(sz), C.POINTER(Info)]
    L.flo_decode_i16_device.restype = C.c_int
    L.flo_decode_i16_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(Info)]
    L.flo_stream_encode_frames.restype = C.c_int
    L.flo_stream_encode_frames.argtypes = [vp, vp, sz, u32, u8, u8, u8, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(u32)]
    L.flo_waveform_peaks.restype = C.c_int
    L.flo_waveform_peaks.argtypes = [vp, vp, sz, u32, u8, u32, C.POINTER(vp), C.POINTER(sz)]
    L.flo_waveform_peaks_device.restype = C.c_int
    L.flo_waveform_peaks_device.argtypes = [vp, vp, sz, u32, u8, u32, vp, sz, C.POINTER(sz)]
    L.flo_waveform_peaks_count.restype = sz
    L.flo_waveform_peaks_count.argtypes = [sz, u32, u8, u32]
    L.flo_integrated_loudness.restype = C.c_int
    L.flo_integrated_loudness.argtypes = [vp, vp, sz, u32, u8, C.POINTER(C.c_double)]
    L.flo_integrated_loudness_device.restype = C.c_int
    L.flo_integrated_loudness_device.argtypes = [vp, vp, sz, u32, u8, C.POINTER(C.c_double)]
    L.flo_host_alloc.restype = vp
    L.flo_host_alloc.argtypes = [sz]
    L.flo_host_free.restype = None
    L.flo_host_free.argtypes = [vp]
    L.flo_free.restype = None
    L.flo_free.argtypes = [vp]
    L.flo_last_error.restype = C.c_char_p
    L.flo_version.restype = C.c_char_p
    L.flo_device_count.restype = C.c_int
    _lib = L
    return L


def last_error() -> str:
    return lib().flo_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise FloError(last_error())
