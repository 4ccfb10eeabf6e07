"""ctypes loader for the CPU oracle (oracle/libflo_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never from flo_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libflo_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "flo_oracle.c")
    hdr = os.path.join(_HERE, "flo_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr, os.path.join(_HERE, "flo_r128.c"), os.path.join(_HERE, "Makefile"))
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libflo_oracle.so"])
    return _SO


class Candidate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("order", C.c_int32), ("k", C.c_int32), ("size", C.c_int64)]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, i32p, i64p, f32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float))
    L.flo_ref_f32_to_i32.restype = C.c_int32
    L.flo_ref_f32_to_i32.argtypes = [C.c_float]
    L.flo_ref_i32_to_f32.restype = C.c_float
    L.flo_ref_i32_to_f32.argtypes = [C.c_int32]
    L.flo_ref_crc32.restype = C.c_uint32
    L.flo_ref_crc32.argtypes = [C.c_void_p, C.c_size_t]
    L.flo_ref_estimate_rice_parameter_i32.restype = C.c_uint8
    L.flo_ref_estimate_rice_parameter_i32.argtypes = [C.c_void_p, C.c_size_t]
    L.flo_ref_rice_encode_i32.restype = C.c_size_t
    L.flo_ref_rice_encode_i32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.POINTER(u8p)]
    L.flo_ref_rice_decode_i32.restype = None
    L.flo_ref_rice_decode_i32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_size_t, C.c_void_p]
    L.flo_ref_fixed_predictor_residuals.restype = None
    L.flo_ref_fixed_predictor_residuals.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    L.flo_ref_autocorr_int.restype = None
    L.flo_ref_autocorr_int.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    L.flo_ref_levinson_durbin_int.restype = C.c_int
    L.flo_ref_levinson_durbin_int.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_uint8)]
    L.flo_ref_calc_residuals_int.restype = None
    L.flo_ref_calc_residuals_int.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint8, C.c_int, C.c_void_p]
    enc_args = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t,
                C.POINTER(u8p), C.POINTER(C.c_size_t)]
    L.flo_ref_encode.restype = C.c_int
    L.flo_ref_encode.argtypes = enc_args
    L.flo_ref_encode_pcm16.restype = C.c_int
    L.flo_ref_encode_pcm16.argtypes = enc_args
    L.flo_ref_encode_frame_i32.restype = C.c_int
    L.flo_ref_encode_frame_i32.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_uint8, C.c_uint32,
                                           C.c_uint8, C.c_uint8, C.POINTER(u8p), C.POINTER(C.c_size_t)]
    L.flo_ref_channel_candidates.restype = C.c_int
    L.flo_ref_channel_candidates.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.POINTER(Candidate)]
    L.flo_ref_parse.restype = C.c_void_p
    L.flo_ref_parse.argtypes = [C.c_void_p, C.c_size_t]
    L.flo_ref_file_free.restype = None
    L.flo_ref_file_free.argtypes = [C.c_void_p]
    for name, rt in [("sample_rate", C.c_uint32), ("channels", C.c_uint8), ("bit_depth", C.c_uint8),
                     ("level", C.c_uint8), ("total_samples", C.c_uint64), ("crc32", C.c_uint32),
                     ("data_offset", C.c_uint64), ("data_size", C.c_uint64), ("meta_size", C.c_uint64),
                     ("num_frames", C.c_uint32)]:
        fn = getattr(L, "flo_ref_file_" + name)
        fn.restype = rt
        fn.argtypes = [C.c_void_p]
    L.flo_ref_file_frame_info.restype = None
    L.flo_ref_file_frame_info.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_uint32)]
    L.flo_ref_file_channel_info.restype = None
    L.flo_ref_file_channel_info.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8),
                                            C.POINTER(C.c_uint64), C.c_void_p]
    L.flo_ref_file_decode_frame_coded.restype = C.c_int
    L.flo_ref_file_decode_frame_coded.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.flo_ref_decode.restype = C.c_int
    L.flo_ref_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(f32p), C.POINTER(C.c_size_t)]
    L.flo_ref_decode_i32.restype = C.c_int
    L.flo_ref_decode_i32.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(i32p), C.POINTER(C.c_size_t)]
    L.flo_ref_free.restype = None
    L.flo_ref_free.argtypes = [C.c_void_p]
    L.flo_ref_kweighting_coeffs.restype = None
    L.flo_ref_kweighting_coeffs.argtypes = [C.c_double, C.POINTER(C.c_double)]
    L.flo_ref_r128_integrated.restype = C.c_double
    L.flo_ref_r128_integrated.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_size_t)]
    L.flo_ref_last_error.restype = C.c_char_p
    L.flo_ref_last_error.argtypes = []
    _lib = L
    return L


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _take(p, n: int) -> bytes:
    out = C.string_at(p, n) if n else b""
    lib().flo_ref_free(p)
    return out


# ---- primitives -----------------------------------------------------------
def f32_to_i32(x: float) -> int:
    return int(lib().flo_ref_f32_to_i32(C.c_float(x)))


def crc32(data: bytes) -> int:
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    return int(lib().flo_ref_crc32(_ptr(buf), len(data)))


def estimate_rice_parameter_i32(r: Sequence[int]) -> int:
    a = np.ascontiguousarray(r, dtype=np.int32)
    return int(lib().flo_ref_estimate_rice_parameter_i32(_ptr(a), a.size))


def rice_encode_i32(r: Sequence[int], k: int) -> bytes:
    a = np.ascontiguousarray(r, dtype=np.int32)
    p = C.POINTER(C.c_uint8)()
    n = lib().flo_ref_rice_encode_i32(_ptr(a), a.size, k, C.byref(p))
    return _take(p, n)


def rice_decode_i32(enc: bytes, k: int, target_len: int) -> np.ndarray:
    buf = np.frombuffer(enc, dtype=np.uint8) if len(enc) else np.zeros(1, np.uint8)
    out = np.zeros(max(target_len, 1), np.int32)
    lib().flo_ref_rice_decode_i32(_ptr(buf), len(enc), k, target_len, _ptr(out))
    return out[:target_len]


def fixed_predictor_residuals(s: Sequence[int], order: int) -> np.ndarray:
    a = np.ascontiguousarray(s, dtype=np.int32)
    out = np.zeros(max(a.size, 1), np.int32)
    lib().flo_ref_fixed_predictor_residuals(_ptr(a), a.size, order, _ptr(out))
    return out[:a.size]


def autocorr_int(s: Sequence[int], order: int) -> np.ndarray:
    a = np.ascontiguousarray(s, dtype=np.int32)
    out = np.zeros(order + 1, np.int64)
    lib().flo_ref_autocorr_int(_ptr(a), a.size, order, _ptr(out))
    return out


def levinson_durbin_int(ac: Sequence[int], order: int):
    a = np.ascontiguousarray(ac, dtype=np.int64)
    co = np.zeros(max(order, 1), np.int32)
    sh = C.c_uint8(0)
    ok = lib().flo_ref_levinson_durbin_int(_ptr(a), order, _ptr(co), C.byref(sh))
    return (co[:order].copy(), int(sh.value)) if ok else None


def calc_residuals_int(s: Sequence[int], coeffs: Sequence[int], shift: int, order: int) -> np.ndarray:
    a = np.ascontiguousarray(s, dtype=np.int32)
    co = np.ascontiguousarray(coeffs, dtype=np.int32)
    out = np.zeros(max(a.size, 1), np.int32)
    lib().flo_ref_calc_residuals_int(_ptr(a), a.size, _ptr(co), shift, order, _ptr(out))
    return out[:a.size]


# ---- encode path ----------------------------------------------------------
def _encode(fn, samples: np.ndarray, sample_rate: int, channels: int, bit_depth: int, level: int,
            metadata: bytes) -> bytes:
    meta = np.frombuffer(metadata, dtype=np.uint8) if len(metadata) else np.zeros(1, np.uint8)
    s = samples if samples.size else np.zeros(1, samples.dtype)
    p = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    rc = fn(_ptr(s), samples.size, sample_rate, channels, bit_depth, level, _ptr(meta), len(metadata),
            C.byref(p), C.byref(n))
    if rc != 0:
        raise RuntimeError(lib().flo_ref_last_error().decode())
    return _take(p, n.value)


def encode(samples, sample_rate: int, channels: int, bit_depth: int = 16, level: int = 5,
           metadata: bytes = b"") -> bytes:
    """Encoder::new(sr, ch, bits).with_compression(level).encode(samples, metadata)."""
    a = np.ascontiguousarray(samples, dtype=np.float32)
    return _encode(lib().flo_ref_encode, a, sample_rate, channels, bit_depth, level, metadata)


def encode_pcm16(pcm, sample_rate: int, channels: int, bit_depth: int = 16, level: int = 5,
                 metadata: bytes = b"") -> bytes:
    a = np.ascontiguousarray(pcm, dtype=np.int16)
    return _encode(lib().flo_ref_encode_pcm16, a, sample_rate, channels, bit_depth, level, metadata)


def encode_frame_i32(channels: List[np.ndarray], frame_samples: int, flags: int, level: int) -> bytes:
    chs = [np.ascontiguousarray(c, dtype=np.int32) for c in channels]
    keep = [c if c.size else np.zeros(1, np.int32) for c in chs]
    ptrs = (C.c_void_p * len(chs))(*[c.ctypes.data for c in keep])
    lens = (C.c_size_t * len(chs))(*[c.size for c in chs])
    p = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    rc = lib().flo_ref_encode_frame_i32(ptrs, lens, len(chs), frame_samples, flags, level, C.byref(p), C.byref(n))
    if rc != 0:
        raise RuntimeError(lib().flo_ref_last_error().decode())
    return _take(p, n.value)


def channel_candidates(s: Sequence[int], level: int) -> List[dict]:
    a = np.ascontiguousarray(s, dtype=np.int32)
    buf = (Candidate * 16)()
    n = lib().flo_ref_channel_candidates(_ptr(a if a.size else np.zeros(1, np.int32)), a.size, level, buf)
    return [dict(kind=buf[i].kind, order=buf[i].order, k=buf[i].k, size=buf[i].size) for i in range(n)]


# ---- reader / decoder -----------------------------------------------------
@dataclass
class ChannelInfo:
    n_coeffs: int
    shift_bits: int
    encoding: int
    k: int
    residual_bytes: int
    coeffs: List[int]


@dataclass
class FrameInfo:
    frame_type: int
    frame_samples: int
    flags: int
    byte_offset: int
    frame_size: int
    timestamp_ms: int
    channels: List[ChannelInfo]


class FloFile:
    """Reader::read (reader.rs:16-52) result."""

    def __init__(self, data: bytes):
        self.data = bytes(data)
        self._buf = np.frombuffer(self.data, dtype=np.uint8)
        self._h = lib().flo_ref_parse(_ptr(self._buf), len(self.data))
        if not self._h:
            raise ValueError(lib().flo_ref_last_error().decode())
        L = lib()
        h = self._h
        self.sample_rate = int(L.flo_ref_file_sample_rate(h))
        self.channels = int(L.flo_ref_file_channels(h))
        self.bit_depth = int(L.flo_ref_file_bit_depth(h))
        self.level = int(L.flo_ref_file_level(h))
        self.total_samples = int(L.flo_ref_file_total_samples(h))
        self.crc32 = int(L.flo_ref_file_crc32(h))
        self.data_offset = int(L.flo_ref_file_data_offset(h))
        self.data_size = int(L.flo_ref_file_data_size(h))
        self.meta_size = int(L.flo_ref_file_meta_size(h))
        self.num_frames = int(L.flo_ref_file_num_frames(h))
        self.frames: List[FrameInfo] = []
        for i in range(self.num_frames):
            t, fl = C.c_uint8(), C.c_uint8()
            ns, fs, ts = C.c_uint32(), C.c_uint32(), C.c_uint32()
            off = C.c_uint64()
            L.flo_ref_file_frame_info(h, i, C.byref(t), C.byref(ns), C.byref(fl), C.byref(off), C.byref(fs), C.byref(ts))
            chans = []
            nch = 1 if t.value == 253 else self.channels
            for c in range(nch):
                nc, rb = C.c_uint32(), C.c_uint64()
                sb, en, k = C.c_uint8(), C.c_uint8(), C.c_uint8()
                co = np.zeros(12, np.int32)
                L.flo_ref_file_channel_info(h, i, c, C.byref(nc), C.byref(sb), C.byref(en), C.byref(k), C.byref(rb), _ptr(co))
                chans.append(ChannelInfo(nc.value, sb.value, en.value, k.value, rb.value, co[:nc.value].tolist()))
            self.frames.append(FrameInfo(t.value, ns.value, fl.value, off.value, fs.value, ts.value, chans))

    def frame_bytes(self, i: int) -> bytes:
        fr = self.frames[i]
        s = self.data_offset + fr.byte_offset
        return self.data[s:s + fr.frame_size]

    def data_chunk(self) -> bytes:
        return self.data[self.data_offset:self.data_offset + self.data_size]

    def decode_frame_coded(self, i: int) -> np.ndarray:
        """decode_channel_int for each channel: [channels, frame_samples] int32 (coded domain)."""
        fr = self.frames[i]
        out = np.zeros((max(len(fr.channels), 1), max(fr.frame_samples, 1)), np.int32)
        out = np.ascontiguousarray(out[:, :fr.frame_samples].reshape(len(fr.channels), fr.frame_samples))
        if out.size:
            lib().flo_ref_file_decode_frame_coded(self._h, i, _ptr(out))
        return out

    def close(self):
        if self._h:
            lib().flo_ref_file_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode(data: bytes) -> np.ndarray:
    buf = np.frombuffer(data, dtype=np.uint8)
    p = C.POINTER(C.c_float)()
    n = C.c_size_t(0)
    if lib().flo_ref_decode(_ptr(buf), len(data), C.byref(p), C.byref(n)) != 0:
        raise ValueError(lib().flo_ref_last_error().decode())
    out = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value].copy()
    lib().flo_ref_free(p)
    return out


def decode_i32(data: bytes) -> np.ndarray:
    buf = np.frombuffer(data, dtype=np.uint8)
    p = C.POINTER(C.c_int32)()
    n = C.c_size_t(0)
    if lib().flo_ref_decode_i32(_ptr(buf), len(data), C.byref(p), C.byref(n)) != 0:
        raise ValueError(lib().flo_ref_last_error().decode())
    out = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value].copy()
    lib().flo_ref_free(p)
    return out


# ---- streaming/encoder.rs ------------------------------------------------------------------------------
def _f64_as_u32(v: float) -> int:
    if v != v or v <= 0.0:
        return 0
    return min(int(v), 0xFFFFFFFF)


class StreamingEncoderRef:
    """Restatement of StreamingEncoder (libflo/src/streaming/encoder.rs:5-257) over this oracle's Encoder and
    Reader.  Test infrastructure only.  Parity unpinned: the reference ships no streaming bitstreams; the
    restatement follows the source line by line and is cross-checked through the oracle's own reader."""

    def __init__(self, sample_rate: int, channels: int, bit_depth: int):
        self.sample_rate, self.channels, self.bit_depth = sample_rate, channels, bit_depth
        self.level = 5                                                        # encoder.rs:39
        self.samples_per_frame = sample_rate                                  # encoder.rs:34
        self.buf: List[float] = []
        self.pending: List[dict] = []
        self.total_samples = 0
        self.frame_index = 0

    def with_compression(self, level: int) -> "StreamingEncoderRef":          # encoder.rs:51-56
        self.level = min(level, 9)
        return self

    def pending_samples(self) -> int:
        return len(self.buf) // self.channels

    def pending_frames(self) -> int:
        return len(self.pending)

    def _serialize_channel(self, ch: "ChannelInfo", body: bytes, ftype: int) -> bytes:   # encoder.rs:243-257
        res = body[len(body) - ch.residual_bytes:] if ch.residual_bytes else b""
        if ftype == 0:
            return b""
        if ftype in (254, 253):
            return body[:ch.residual_bytes]
        out = bytes([ch.k])
        for c in ch.coeffs:
            out += int(c).to_bytes(4, "little", signed=True)
        return out + res

    def _encode_frame_data(self, samples: np.ndarray) -> bytes:               # encoder.rs:216-241
        f = FloFile(encode(samples, self.sample_rate, self.channels, self.bit_depth, self.level, b""))
        if not f.frames:
            raise ValueError("No frames encoded")
        fr = f.frames[0]
        raw = f.frame_bytes(0)
        out = bytes([fr.frame_type]) + int(fr.frame_samples).to_bytes(4, "little") + bytes([fr.flags])
        pos = 6
        for ch in fr.channels:
            size = int.from_bytes(raw[pos:pos + 4], "little")
            body = raw[pos + 4:pos + 4 + size]
            pos += 4 + size
            cd = self._serialize_channel(ch, body, fr.frame_type)
            out += len(cd).to_bytes(4, "little") + cd
        return out

    def push_samples(self, samples) -> None:                                  # encoder.rs:71-75, 189-213
        self.buf.extend(np.asarray(samples, np.float32).reshape(-1).tolist())
        frame_samples = self.samples_per_frame * self.channels
        while len(self.buf) >= frame_samples:
            frame = np.asarray(self.buf[:frame_samples], np.float32)
            del self.buf[:frame_samples]
            ts = _f64_as_u32(self.total_samples / float(self.sample_rate) * 1000.0)
            self.pending.append({"index": self.frame_index, "timestamp_ms": ts, "data": self._encode_frame_data(frame),
                                 "samples": self.samples_per_frame})
            self.total_samples += self.samples_per_frame
            self.frame_index += 1

    def next_frame(self):                                                     # encoder.rs:78-84
        return self.pending.pop(0) if self.pending else None

    def flush(self):                                                          # encoder.rs:87-110
        if not self.buf:
            return None
        per = len(self.buf) // self.channels
        ts = _f64_as_u32(self.total_samples / float(self.sample_rate) * 1000.0)
        fr = {"index": self.frame_index, "timestamp_ms": ts, "data": self._encode_frame_data(np.asarray(self.buf, np.float32)),
              "samples": per}
        self.total_samples += per
        self.frame_index += 1
        self.buf = []
        return fr

    def finalize(self, metadata: bytes = b"") -> bytes:                       # encoder.rs:113-183
        fr = self.flush()
        if fr is not None:
            self.pending.append(fr)
        toc = len(self.pending).to_bytes(4, "little")
        off = 0
        for f in self.pending:
            toc += f["index"].to_bytes(4, "little") + off.to_bytes(8, "little") + len(f["data"]).to_bytes(4, "little") \
                + f["timestamp_ms"].to_bytes(4, "little")
            off += len(f["data"])
        data = b"".join(f["data"] for f in self.pending)
        total = sum(f["samples"] for f in self.pending)
        out = b"FLO!" + bytes([1, 2]) + (0).to_bytes(2, "little") + self.sample_rate.to_bytes(4, "little")
        out += bytes([self.channels, self.bit_depth]) + total.to_bytes(8, "little") + bytes([self.level, 0, 0, 0])
        out += crc32(data).to_bytes(4, "little")
        for v in (66, len(toc), len(data), 0, len(metadata)):
            out += v.to_bytes(8, "little")
        out += toc + data + bytes(metadata)
        self.pending = []
        return out


# ---- EBU R128 integrated loudness (ebu_r128.rs:182-313; flo_r128.c) -- parity unpinned ------------------------
def kweighting_coeffs(sample_rate: float) -> List[float]:
    out = (C.c_double * 10)()
    lib().flo_ref_kweighting_coeffs(float(sample_rate), out)
    return list(out)


def r128_integrated_lufs(samples, channels: int, sample_rate: int, blocks: bool = False):
    """compute_ebu_r128_loudness(samples, channels, sample_rate).integrated_lufs; with blocks=True also the 400 ms
    block energies the gating runs over."""
    x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
    bp, nb = C.POINTER(C.c_double)(), C.c_size_t()
    v = lib().flo_ref_r128_integrated(_ptr(x), x.size, channels, sample_rate, C.byref(bp) if blocks else None, C.byref(nb) if blocks else None)
    if not blocks:
        return float(v)
    be = np.ctypeslib.as_array(bp, shape=(nb.value,)).copy() if nb.value else np.zeros(0)
    if bp:
        lib().flo_ref_free(C.cast(bp, C.c_void_p))
    return float(v), be


# ---- reflo's other ingest arms (reflo/src/audio.rs:255-269), numpy f32 arithmetic --------------------------
def s32_to_f32(pcm) -> np.ndarray:
    """S32 arm: `s as f32 * (1.0 / 2147483648.0)` (audio.rs:255-262)."""
    return np.asarray(pcm, np.int32).astype(np.float32) * np.float32(1.0 / 2147483648.0)


def u8_to_f32(pcm) -> np.ndarray:
    """U8 arm: `(s as f32 - 128.0) / 128.0` (audio.rs:263-269)."""
    return (np.asarray(pcm, np.uint8).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
