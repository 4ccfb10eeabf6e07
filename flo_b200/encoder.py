"""Host-side mirror of the reference's lossless encoder interface over the C ABI.

    Encoder(sample_rate, channels, bit_depth).with_compression(level).encode(samples, metadata) -> bytes

is `libflo_audio::Encoder::{new, with_compression, encode}` (libflo/src/lossless/encoder.rs:17-45,
Docs/rust-api.md:44-73) with the same argument meaning and the same bytes out.  Everything is
computed by the CUDA kernels behind include/flo_b200.h; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import FMT_F32, FMT_PCM16, FMT_S32, FMT_U8, FloError


class Context:
    """One GPU + its streams and scratch arenas (flo_ctx)."""

    def __init__(self, device: int = 0):
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.flo_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.flo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ---------------------------------------------------------------
    def set_stream(self, cuda_stream: int) -> None:
        _lib.check(self._L.flo_ctx_set_stream(self._h, C.c_void_p(int(cuda_stream) or None)))

    def enable_report(self, on: bool = True) -> None:
        _lib.check(self._L.flo_ctx_enable_report(self._h, int(on)))

    def read_report(self, frame: int, channel: int) -> List[Tuple[int, int]]:
        """[(k, size)] for candidates raw, fixed 0..4, lpc 5..12 (size -1 = absent)."""
        buf = (_lib.CandReport * 14)()
        _lib.check(self._L.flo_ctx_read_report(self._h, frame, channel, buf))
        return [(int(b.k), int(b.size)) for b in buf]

    def last_timing(self) -> dict:
        ms = (C.c_float * 6)()
        n = C.c_uint32()
        _lib.check(self._L.flo_ctx_last_timing(self._h, ms, C.byref(n)))
        keys = ["device_ms", "encode_ms", "crc_ms", "misc_ms", "h2d_ms", "d2h_ms"]
        d = {k: float(v) for k, v in zip(keys, ms)}
        d["launches"] = int(n.value)
        return d

    def last_counters(self) -> dict:
        buf = (C.c_uint64 * 24)()
        _lib.check(self._L.flo_ctx_last_counters(self._h, buf))
        keys = ["loud_frames", "exact_rounds", "lpc_window_hits", "lpc_window_misses", "fixed_exact", "pruned"]
        d = {k: int(v) for k, v in zip(keys, buf)}
        d["phase_clocks"] = {k: int(buf[8 + i]) for i, k in enumerate(["ingest", "analysis", "lookback", "pack", "frame", "pack_codes", "pack_scan", "pack_emit",
                                                                 "pass1", "levinson", "pass2", "exact_select", "ingest_desc", "ingest_thread0", "pack_flush_thread0", "deferred_wait"])}
        return d

    # -- encode entries ----------------------------------------------------------------
    def encode_batch(self, tracks: Sequence["TrackSpec"], level: int = 5, fmt: int = FMT_F32, views: bool = False):
        """Returns a list of `bytes`; with views=True a BatchResult of zero-copy uint8 arrays over the library's
        (pinned) output buffers, valid until its close()."""
        n = len(tracks)
        if n == 0:
            return []
        dt = {FMT_F32: np.float32, FMT_PCM16: np.int16, FMT_U8: np.uint8, FMT_S32: np.int32}[fmt]
        # the flo_track table is filled column by column (a batch can hold thousands of tracks)
        ss = [np.ascontiguousarray(t.samples, dtype=dt).reshape(-1) for t in tracks]
        metas = [bytes(t.metadata) if t.metadata else b"" for t in tracks]
        mbufs = [C.create_string_buffer(m, len(m)) if m else None for m in metas]
        keep = (ss, mbufs)
        rec = np.zeros(n, dtype=_TRACK_DT)
        rec["samples"] = [s.__array_interface__["data"][0] if s.size else 0 for s in ss]
        rec["n_interleaved"] = [s.size for s in ss]
        rec["sample_rate"] = [_u(t.sample_rate, 32, "sample_rate") for t in tracks]
        rec["channels"] = [_u(t.channels, 8, "channels") for t in tracks]
        rec["bit_depth"] = [_u(t.bit_depth, 8, "bit_depth") for t in tracks]
        rec["meta"] = [C.addressof(b) if b is not None else 0 for b in mbufs]
        rec["meta_len"] = [len(m) for m in metas]
        arr = (_lib.Track * n).from_buffer(rec)
        outs = (_lib.Out * n)()
        _lib.check(self._L.flo_encode_batch(self._h, arr, n, fmt, min(int(level), 255), outs))
        if views:
            return BatchResult(self._L, outs)
        res = []
        for o in outs:
            res.append(C.string_at(o.data, o.len) if o.len else b"")
            self._L.flo_free(o.data)
        return res

    def encode_batch_device(self, dev_ptrs: Sequence[int], n_interleaved: Sequence[int], sample_rate: Sequence[int],
                            channels: Sequence[int], bit_depth: Sequence[int], d_out: int, d_out_capacity: int,
                            level: int = 5, fmt: int = FMT_F32, metadata: Optional[Sequence[bytes]] = None):
        """Device-resident batch: inputs and the output arena stay in HBM.  Returns (offsets, lens)."""
        n = len(dev_ptrs)
        arr = (_lib.Track * n)()
        keep = []
        for i in range(n):
            m = bytes(metadata[i]) if metadata is not None and metadata[i] else b""
            mb = C.create_string_buffer(m, len(m)) if m else None
            keep.append(mb)
            arr[i].samples = int(dev_ptrs[i]) or None
            arr[i].n_interleaved = int(n_interleaved[i])
            arr[i].sample_rate = int(sample_rate[i])
            arr[i].channels = int(channels[i])
            arr[i].bit_depth = int(bit_depth[i])
            arr[i].meta = C.addressof(mb) if mb is not None else None
            arr[i].meta_len = len(m)
        off = np.zeros(n, np.uint64)
        ln = np.zeros(n, np.uint64)
        _lib.check(self._L.flo_encode_batch_device(
            self._h, arr, n, fmt, min(int(level), 255), C.c_void_p(int(d_out)), int(d_out_capacity),
            off.ctypes.data_as(C.POINTER(C.c_uint64)), ln.ctypes.data_as(C.POINTER(C.c_uint64))))
        return off, ln

    # -- lossless decoder (include/flo_b200.h: flo_decode / flo_decode_device) ---------
    def decode(self, data) -> Tuple[np.ndarray, dict]:
        """Decoder::decode (libflo/src/lossless/decoder.rs:14-18): interleaved f32 samples + header info.
        The array is a zero-copy view of the library's (pinned) result block, released with the array."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        out, n, info = C.c_void_p(), C.c_size_t(), _lib.Info()
        ptr = buf.ctypes.data if buf.size else C.addressof(C.create_string_buffer(1))
        _lib.check(self._L.flo_decode(self._h, C.c_void_p(ptr), buf.size, C.byref(out), C.byref(n), C.byref(info)))
        if n.value == 0:
            self._L.flo_free(out)
            return np.zeros(0, dtype=np.float32), _info_dict(info)
        raw = (C.c_float * n.value).from_address(out.value)
        weakref.finalize(raw, self._L.flo_free, C.c_void_p(out.value))
        return np.frombuffer(raw, dtype=np.float32), _info_dict(info)

    def decode_i16(self, data) -> Tuple[np.ndarray, dict]:
        """flo_decode_i16: the decoder's integer samples saturated to int16 (half the bytes of decode())."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        out, n, info = C.c_void_p(), C.c_size_t(), _lib.Info()
        ptr = buf.ctypes.data if buf.size else C.addressof(C.create_string_buffer(1))
        _lib.check(self._L.flo_decode_i16(self._h, C.c_void_p(ptr), buf.size, C.byref(out), C.byref(n), C.byref(info)))
        if n.value == 0:
            self._L.flo_free(out)
            return np.zeros(0, dtype=np.int16), _info_dict(info)
        raw = (C.c_int16 * n.value).from_address(out.value)
        weakref.finalize(raw, self._L.flo_free, C.c_void_p(out.value))
        return np.frombuffer(raw, dtype=np.int16), _info_dict(info)

    def decode_i16_device(self, d_file: int, length: int, d_out: int, capacity: int) -> Tuple[int, dict]:
        """Device-resident flo_decode_i16 (capacity in int16 samples)."""
        n, info = C.c_size_t(), _lib.Info()
        _lib.check(self._L.flo_decode_i16_device(self._h, C.c_void_p(int(d_file)), int(length), C.c_void_p(int(d_out) or None),
                                                 int(capacity), C.byref(n), C.byref(info)))
        return int(n.value), _info_dict(info)

    def decode_device(self, d_file: int, length: int, d_out: int, capacity: int) -> Tuple[int, dict]:
        """Device-resident decode: `d_file` / `d_out` are device pointers (capacity in floats).
        Returns (interleaved sample count, info)."""
        n, info = C.c_size_t(), _lib.Info()
        _lib.check(self._L.flo_decode_device(self._h, C.c_void_p(int(d_file)), int(length), C.c_void_p(int(d_out) or None),
                                             int(capacity), C.byref(n), C.byref(info)))
        return int(n.value), _info_dict(info)

    def output_bound(self, n_interleaved: Sequence[int], sample_rate: Sequence[int], channels: Sequence[int],
                     meta_len: Optional[Sequence[int]] = None) -> int:
        n = len(n_interleaved)
        arr = (_lib.Track * n)()
        for i in range(n):
            arr[i].n_interleaved = int(n_interleaved[i])
            arr[i].sample_rate = int(sample_rate[i])
            arr[i].channels = int(channels[i])
            arr[i].bit_depth = 16
            arr[i].samples = 16 if n_interleaved[i] else None       # never dereferenced by the bound
            arr[i].meta_len = 0
        extra = int(sum(meta_len)) if meta_len is not None else 0
        b = int(self._L.flo_output_bound(arr, n))
        if b == 0:
            raise FloError(_lib.last_error())
        return b + extra


def _info_dict(info) -> dict:
    return {k: int(getattr(info, k)) for k, _ in _lib.Info._fields_}


# numpy view of flo_track (include/flo_b200.h): the same offsets as the ctypes structure
_TRACK_DT = np.dtype({"names": [f[0] for f in _lib.Track._fields_],
                      "formats": [np.uint64, np.uint64, np.uint32, np.uint8, np.uint8, np.uint64, np.uint64],
                      "offsets": [getattr(_lib.Track, f[0]).offset for f in _lib.Track._fields_],
                      "itemsize": C.sizeof(_lib.Track)})


class BatchResult:
    """Zero-copy views of the file images returned by flo_encode_batch; close() hands them back (flo_free).
    The views are made on first access (a batch of thousands of tracks should not pay for arrays nobody reads)."""

    def __init__(self, L, outs):
        self._L, self._outs = L, outs
        self._arrays = None
        # (data, len) pairs as one uint64 table over the ctypes array
        self._table = np.frombuffer(outs, dtype=np.uint64).reshape(-1, 2) if len(outs) else np.zeros((0, 2), np.uint64)

    @property
    def arrays(self):
        if self._arrays is None:
            self._arrays = [np.ctypeslib.as_array(C.cast(o.data, C.POINTER(C.c_uint8)), shape=(o.len,)) if o.len
                            else np.zeros(0, np.uint8) for o in (self._outs or [])]
        return self._arrays

    def __len__(self):
        return int(self._table.shape[0])

    def __getitem__(self, i):
        return self.arrays[i]

    def total_bytes(self) -> int:
        return int(self._table[:, 1].sum())

    def close(self) -> None:
        if self._outs is not None:
            self._arrays = []
            free = self._L.flo_free
            for ptr in self._table[:, 0].tolist():
                free(ptr)
            self._outs = None
            self._table = np.zeros((0, 2), np.uint64)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _u(v: int, bits: int, name: str) -> int:
    v = int(v)
    if v < 0 or v >= (1 << bits):
        raise FloError(f"{name}={v} does not fit u{bits}")
    return v


@dataclass
class TrackSpec:
    samples: np.ndarray            # interleaved f32 (or int16 for the PCM16 entry)
    sample_rate: int
    channels: int
    bit_depth: int = 16
    metadata: bytes = b""


_ctx_lock = threading.Lock()
_ctxs: dict = {}


def default_context(device: int = 0) -> Context:
    with _ctx_lock:
        c = _ctxs.get(device)
        if c is None:
            c = _ctxs[device] = Context(device)
        return c


class Encoder:
    """libflo_audio::Encoder (libflo/src/lossless/encoder.rs:9-45)."""

    def __init__(self, sample_rate: int = 44100, channels: int = 1, bit_depth: int = 16, *, device: int = 0,
                 context: Optional[Context] = None):
        # Default = (44100, 1, 16): Docs/rust-api.md, impl Default for Encoder
        self.sample_rate = _u(sample_rate, 32, "sample_rate")
        self.channels = _u(channels, 8, "channels")
        self.bit_depth = _u(bit_depth, 8, "bit_depth")
        self.compression_level = 5                          # encoder.rs:22
        self._device = device
        self._ctx = context

    def with_compression(self, level: int) -> "Encoder":
        self.compression_level = min(_u(level, 8, "level"), 9)   # encoder.rs:26-29
        return self

    def _context(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context(self._device)
        return self._ctx

    def encode(self, samples, metadata: bytes = b"") -> bytes:
        """Encoder::encode(&self, samples: &[f32], metadata: &[u8]) -> FloResult<Vec<u8>> (encoder.rs:32-45)."""
        t = TrackSpec(np.asarray(samples, dtype=np.float32), self.sample_rate, self.channels, self.bit_depth, metadata)
        return self._context().encode_batch([t], self.compression_level, FMT_F32)[0]

    def encode_pcm(self, pcm, metadata: bytes = b"") -> bytes:
        """reflo's integer ingest arms (reflo/src/audio.rs:247-269) + Encoder::encode, on the device: the array's
        dtype picks the arm (uint8, int16 or int32)."""
        a = np.asarray(pcm)
        fmt = {np.dtype(np.uint8): FMT_U8, np.dtype(np.int16): FMT_PCM16, np.dtype(np.int32): FMT_S32}.get(a.dtype)
        if fmt is None:
            raise FloError(f"unsupported PCM dtype {a.dtype}")
        t = TrackSpec(a, self.sample_rate, self.channels, self.bit_depth, metadata)
        return self._context().encode_batch([t], self.compression_level, fmt)[0]

    def encode_pcm16(self, pcm, metadata: bytes = b"") -> bytes:
        """reflo's S16 ingest (reflo/src/audio.rs:247-254) + Encoder::encode, fused on the device."""
        t = TrackSpec(np.asarray(pcm, dtype=np.int16), self.sample_rate, self.channels, self.bit_depth, metadata)
        return self._context().encode_batch([t], self.compression_level, FMT_PCM16)[0]


def encode_batch(tracks: Sequence[TrackSpec], level: int = 5, fmt: int = FMT_F32, device: int = 0) -> List[bytes]:
    """Loop of Encoder::encode over tracks (reflo/src/main.rs:218-276) as one device pass."""
    return default_context(device).encode_batch(tracks, min(int(level), 9), fmt)


class Decoder:
    """libflo_audio::Decoder (libflo/src/lossless/decoder.rs:6-18): `Decoder().decode(data) -> samples`."""

    def __init__(self, *, device: int = 0, context: Optional[Context] = None):
        self._device = device
        self._ctx = context

    def _context(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context(self._device)
        return self._ctx

    def decode(self, data) -> np.ndarray:
        """Decoder::decode(&self, data: &[u8]) -> FloResult<Vec<f32>>: interleaved f32 samples."""
        return self._context().decode(data)[0]

    def decode_i16(self, data) -> np.ndarray:
        """The same samples as 16-bit integers (before the reference's final i32 -> f32 conversion, saturated)."""
        return self._context().decode_i16(data)[0]

    def decode_with_info(self, data) -> Tuple[np.ndarray, dict]:
        return self._context().decode(data)
