"""Streaming encoder semantics on top of the batch encoder.

Mirror of `libflo_audio::StreamingEncoder` (libflo/src/streaming/encoder.rs:5-257): `push_samples` /
`next_frame` / `flush` / `finalize`.  The reference encodes every 1-second frame with `Encoder::encode`, parses
the one-frame file back and re-serialises each channel in its own, Writer-incompatible layout
(`serialize_channel`, encoder.rs:243-257: `[rice_parameter][coeffs as i32 LE ...][residual bytes]`, no order /
shift / encoding bytes).  Here all complete frames of one `push_samples` call go through ONE device pass
(`flo_encode_batch` with one single-frame track per frame); the re-serialisation is byte shuffling on the host.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from ._lib import FMT_F32, FloError
from .encoder import Context, TrackSpec, _u, default_context


@dataclass
class EncodedFrame:
    """streaming/encoder.rs:18-29"""
    index: int
    timestamp_ms: int
    data: bytes
    samples: int


def _f64_as_u32(v: float) -> int:
    """Rust `f64 as u32`: truncating, saturating, NaN -> 0."""
    if v != v or v <= 0.0:
        return 0
    return min(int(v), 0xFFFFFFFF)


def reserialize_frame(flo: bytes, channels: int) -> bytes:
    """encode_frame_data's second half (encoder.rs:216-241): frame 0 of a one-frame file image, re-serialised.
    Reads exactly what Reader::read_channel_data reads (reader.rs:168-247) for the frame types the encoder writes."""
    if len(flo) < 74 or struct.unpack_from("<I", flo, 70)[0] == 0:
        raise FloError("No frames encoded")                                   # encoder.rs:222-224
    pos = 74 + 20 * struct.unpack_from("<I", flo, 70)[0]
    ftype, n, flags = flo[pos], struct.unpack_from("<I", flo, pos + 1)[0], flo[pos + 5]
    out = bytearray(flo[pos:pos + 6])
    pos += 6
    for _ in range(channels):
        size = struct.unpack_from("<I", flo, pos)[0]
        body = flo[pos + 4:pos + 4 + size]
        pos += 4 + size
        if ftype == 0:                                                        # Silence: nothing
            ch = b""
        elif ftype == 254:                                                    # Raw: at most 2 * frame_samples bytes
            ch = body[:min(2 * n, size)]
        elif 1 <= ftype <= 12:                                                # ALPC: k, coefficients, residual bytes
            order = body[0]
            p = 1 + 4 * order
            enc = body[p + 1]
            k = body[p + 2] if enc == 0 else 0
            p += 3 if enc == 0 else 2
            ch = bytes([k]) + body[1:1 + 4 * order] + body[p:]
        else:
            ch = b""
        out += struct.pack("<I", len(ch)) + ch
    return bytes(out)


class StreamingEncoder:
    """libflo_audio::StreamingEncoder (streaming/encoder.rs:5-257)."""

    def __init__(self, sample_rate: int, channels: int, bit_depth: int, *, device: int = 0, context: Optional[Context] = None):
        self.sample_rate = _u(sample_rate, 32, "sample_rate")
        self.channels = _u(channels, 8, "channels")
        self.bit_depth = _u(bit_depth, 8, "bit_depth")
        self.compression_level = 5                                            # encoder.rs:39
        self.samples_per_frame = self.sample_rate                             # encoder.rs:34
        self._buf = np.zeros(0, np.float32)
        self._pending: List[EncodedFrame] = []
        self._total_samples = 0
        self._frame_index = 0
        self._device = device
        self._ctx = context

    def with_compression(self, level: int) -> "StreamingEncoder":
        self.compression_level = min(_u(level, 8, "level"), 9)                # encoder.rs:51-56
        return self

    def _context(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context(self._device)
        return self._ctx

    def _per_channel(self, n_interleaved: int) -> int:
        if self.channels == 0:
            raise FloError("channels must be non-zero")                      # the reference divides by zero here
        return n_interleaved // self.channels

    def pending_samples(self) -> int:
        return self._per_channel(self._buf.size)                              # encoder.rs:59-61

    def pending_frames(self) -> int:
        return len(self._pending)

    def _encode_run(self, samples: np.ndarray) -> List[bytes]:
        """encode_frame_data (encoder.rs:216-241) for every frame of a contiguous run of samples: one device pass
        and the re-serialisation in C (flo_stream_encode_frames, include/flo_b200.h)."""
        if self.channels == 0 or self.sample_rate == 0:
            raise FloError("channels and sample_rate must be non-zero")
        import ctypes as C
        from . import _lib
        ctx = self._context()
        x = np.ascontiguousarray(samples, dtype=np.float32)
        out, out_len, offs, nfr = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_uint32()
        _lib.check(ctx._L.flo_stream_encode_frames(ctx._h, x.ctypes.data_as(C.c_void_p), x.size, self.sample_rate, self.channels,
                                                   self.bit_depth, self.compression_level, C.byref(out), C.byref(out_len),
                                                   C.byref(offs), C.byref(nfr)))
        try:
            blob = C.string_at(out.value, out_len.value)
            off = np.ctypeslib.as_array(C.cast(offs.value, C.POINTER(C.c_uint64)), shape=(nfr.value + 1,)).copy()
        finally:
            ctx._L.flo_free(out)
            ctx._L.flo_free(offs)
        return [blob[int(off[i]):int(off[i + 1])] for i in range(nfr.value)]

    def push_samples(self, samples) -> None:
        """encoder.rs:71-75 + try_encode_frames (encoder.rs:189-213)."""
        x = np.asarray(samples, dtype=np.float32).reshape(-1)
        self._buf = x if self._buf.size == 0 else np.concatenate([self._buf, x])     # (no copy for the common bulk push)
        frame_samples = self.samples_per_frame * self.channels
        if frame_samples == 0:
            raise FloError("channels and sample_rate must be non-zero")      # the reference loops forever / divides by zero
        nfull = self._buf.size // frame_samples
        if nfull == 0:
            self._buf = self._buf.copy() if self._buf is x else self._buf         # never keep a view of the caller's array
            return
        for data in self._encode_run(self._buf[:nfull * frame_samples]):
            ts = _f64_as_u32(self._total_samples / float(self.sample_rate) * 1000.0)
            self._pending.append(EncodedFrame(self._frame_index, ts, data, self.samples_per_frame))
            self._total_samples += self.samples_per_frame
            self._frame_index = (self._frame_index + 1) & 0xFFFFFFFF
        self._buf = self._buf[nfull * frame_samples:].copy()

    def next_frame(self) -> Optional[EncodedFrame]:
        return self._pending.pop(0) if self._pending else None                # encoder.rs:78-84

    def flush(self) -> Optional[EncodedFrame]:
        """encoder.rs:87-110: the rest of the buffer as one (possibly partial) frame."""
        if self._buf.size == 0:
            return None
        per = self._per_channel(self._buf.size)
        ts = _f64_as_u32(self._total_samples / float(self.sample_rate) * 1000.0)
        run = self._encode_run(self._buf)
        if not run:
            raise FloError("No frames encoded")                               # encoder.rs:222-224
        data = run[0]
        fr = EncodedFrame(self._frame_index, ts, data, per)
        self._total_samples += per
        self._frame_index = (self._frame_index + 1) & 0xFFFFFFFF
        self._buf = np.zeros(0, np.float32)
        return fr

    def finalize(self, metadata: bytes = b"") -> bytes:
        """encoder.rs:113-183: a file image (format version 1.2) from the frames still pending."""
        fr = self.flush()
        if fr is not None:
            self._pending.append(fr)
        toc = bytearray(struct.pack("<I", len(self._pending)))
        off = 0
        for f in self._pending:
            toc += struct.pack("<IQII", f.index, off, len(f.data), f.timestamp_ms)
            off += len(f.data)
        data = b"".join(f.data for f in self._pending)
        total = sum(f.samples for f in self._pending)
        out = bytearray(b"FLO!")
        out += bytes([1, 2]) + struct.pack("<H", 0) + struct.pack("<I", self.sample_rate) + bytes([self.channels, self.bit_depth])
        out += struct.pack("<Q", total) + bytes([self.compression_level, 0, 0, 0]) + struct.pack("<I", zlib.crc32(data) & 0xFFFFFFFF)
        out += struct.pack("<QQQQQ", 66, len(toc), len(data), 0, len(metadata))
        out += toc + data + bytes(metadata)
        self._pending = []
        return bytes(out)
