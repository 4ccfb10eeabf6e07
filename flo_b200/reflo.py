"""reflo-compatible front of the encode path.

`reflo::encode_from_samples` (reflo/src/lib.rs:202-309) builds a `FloMetadata`, serialises it with
rmp-serde's named-map MessagePack (`to_vec_named`, libflo/src/core/metadata.rs:709-711) and hands it to
`Encoder::encode` as the opaque metadata tail.  This module restates that front for the lossless case so
that whole files equal `reflo encode` output byte for byte -- given the wall-clock stamp reflo would have
written (`encoding_time`, reflo/src/lib.rs:257), which the caller supplies.

Only the scalar tag fields (Option<String> / Option<u32> / Option<u64>) are supported; structured ones
(pictures, comments, lyrics, analysis data ...) are refused rather than guessed.
"""
from __future__ import annotations

import struct
from typing import Dict, Optional, Union

import numpy as np

from ._lib import FloError
from .encoder import Context, Encoder

REFLO_VERSION = "reflo 0.1.2"          # CARGO_PKG_VERSION of the reference checkout (reflo/Cargo.toml)

# FloMetadata's scalar fields in declaration order (libflo/src/core/metadata.rs:328-...); serde emits a
# named map in this order and skips None (`skip_serializing_if = "Option::is_none"`).
_S, _U = "str", "uint"
SCALAR_FIELDS = [
    ("title", _S), ("subtitle", _S), ("content_group", _S), ("album", _S), ("original_album", _S), ("set_subtitle", _S),
    ("track_number", _U), ("track_total", _U), ("disc_number", _U), ("disc_total", _U), ("isrc", _S), ("artist", _S),
    ("album_artist", _S), ("conductor", _S), ("remixer", _S), ("original_artist", _S), ("composer", _S), ("lyricist", _S),
    ("original_lyricist", _S), ("encoded_by", _S), ("genre", _S), ("mood", _S), ("bpm", _U), ("key", _S), ("language", _S),
    ("length_ms", _U), ("year", _U), ("recording_time", _S), ("release_time", _S), ("original_release_time", _S),
    ("encoding_time", _S), ("tagging_time", _S), ("copyright", _S), ("produced_notice", _S), ("publisher", _S),
    ("file_owner", _S), ("radio_station", _S), ("radio_station_owner", _S), ("album_sort", _S), ("artist_sort", _S),
    ("title_sort", _S), ("original_filename", _S), ("playlist_delay", _U), ("encoder_settings", _S), ("url_commercial", _S),
    ("url_copyright", _S), ("url_audio_file", _S), ("url_artist", _S), ("url_audio_source", _S), ("url_radio_station", _S),
    ("url_payment", _S), ("url_publisher", _S), ("play_count", _U), ("flo_encoder_version", _S), ("source_format", _S),
]
_KIND = dict(SCALAR_FIELDS)


def _mp_str(s: str) -> bytes:
    b = s.encode("utf-8")
    n = len(b)
    if n < 32:
        return bytes([0xA0 | n]) + b
    if n < 256:
        return bytes([0xD9, n]) + b
    if n < 65536:
        return b"\xda" + struct.pack(">H", n) + b
    return b"\xdb" + struct.pack(">I", n) + b


def _mp_uint(v: int) -> bytes:
    if v < 0:
        raise FloError("metadata integers are unsigned")
    if v < 128:
        return bytes([v])
    if v < 256:
        return bytes([0xCC, v])
    if v < 65536:
        return b"\xcd" + struct.pack(">H", v)
    if v < (1 << 32):
        return b"\xce" + struct.pack(">I", v)
    return b"\xcf" + struct.pack(">Q", v)


def metadata_to_msgpack(fields: Dict[str, Union[str, int, None]]) -> bytes:
    """FloMetadata::to_msgpack for a metadata value that only has scalar fields set."""
    unknown = [k for k in fields if k not in _KIND]
    if unknown:
        raise FloError(f"unsupported metadata fields {unknown}: only scalar FloMetadata fields are supported")
    items = [(name, fields[name]) for name, _ in SCALAR_FIELDS if fields.get(name) is not None]
    n = len(items)
    out = bytearray(bytes([0x80 | n]) if n < 16 else b"\xde" + struct.pack(">H", n))
    for name, v in items:
        out += _mp_str(name)
        out += _mp_str(str(v)) if _KIND[name] == _S else _mp_uint(int(v))
    return bytes(out)


def reflo_metadata(n_interleaved: int, sample_rate: int, channels: int, level: int, encoding_time: str,
                   source_format: Optional[str] = None, original_filename: Optional[str] = None,
                   tags: Optional[Dict[str, Union[str, int]]] = None, version: str = REFLO_VERSION) -> bytes:
    """The metadata bytes `reflo encode` writes for a lossless encode (reflo/src/lib.rs:210-283)."""
    m: Dict[str, Union[str, int, None]] = dict(tags or {})
    m["flo_encoder_version"] = version                                   # lib.rs:247
    m["encoding_time"] = encoding_time                                   # lib.rs:257-259 (wall clock in the reference)
    if source_format is not None:
        m["source_format"] = source_format                               # lib.rs:260
    if original_filename is not None:
        m["original_filename"] = original_filename                       # lib.rs:261
    m["encoder_settings"] = f"Lossless, level {level}"                   # lib.rs:264-273
    total = n_interleaved // channels                                    # lib.rs:276
    m["length_ms"] = int(total / float(sample_rate) * 1000.0)            # lib.rs:277 (f64, truncating cast)
    return metadata_to_msgpack(m)


def encode_from_samples(samples, sample_rate: int, channels: int, level: int = 5, *, encoding_time: str,
                        source_format: Optional[str] = None, original_filename: Optional[str] = None,
                        tags: Optional[Dict[str, Union[str, int]]] = None, context: Optional[Context] = None) -> bytes:
    """reflo::encode_from_samples, lossless arm (reflo/src/lib.rs:202-309): metadata + Encoder::new(sr, ch, 16)
    .with_compression(level).encode(samples, &metadata)."""
    x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
    meta = reflo_metadata(x.size, sample_rate, channels, level, encoding_time, source_format, original_filename, tags)
    return Encoder(sample_rate, channels, 16, context=context).with_compression(level).encode(x, meta)


# ---- WAV front (first half): RIFF parsing on the host, sample conversion on the device --------
def parse_wav(data: bytes):
    """Minimal RIFF/WAVE reader -> (interleaved numpy array, sample_rate, channels).  The array keeps the file's
    sample type (uint8 / int16 / int32 / float32), which selects reflo's ingest arm (reflo/src/audio.rs:238-270).
    24-bit PCM and compressed formats are refused."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise FloError("not a RIFF/WAVE file")
    pos, fmt, body = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        chunk = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = chunk
        elif cid == b"data":
            body = chunk
            break
        pos += 8 + size + (size & 1)
    if fmt is None or body is None or len(fmt) < 16:
        raise FloError("WAVE file without fmt/data chunk")
    tag, channels, sample_rate, _, _, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == 0xFFFE and len(fmt) >= 26:                                      # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
        tag = struct.unpack_from("<H", fmt, 24)[0]
    dt = {(1, 8): np.uint8, (1, 16): np.dtype("<i2"), (1, 32): np.dtype("<i4"), (3, 32): np.dtype("<f4")}.get((tag, bits))
    if dt is None or channels == 0:
        raise FloError(f"unsupported WAVE format (tag {tag}, {bits} bits, {channels} channels)")
    isz = np.dtype(dt).itemsize
    n = len(body) // (isz * channels) * channels
    return np.frombuffer(body, dtype=dt, count=n), int(sample_rate), int(channels)


def encode_wav(wav: bytes, level: int = 5, *, encoding_time: str, source_format: Optional[str] = "WAV",
               original_filename: Optional[str] = None, tags: Optional[Dict[str, Union[str, int]]] = None,
               context: Optional[Context] = None) -> bytes:
    """`reflo encode in.wav` for PCM WAV input (reflo/src/lib.rs:202-309 after reflo/src/audio.rs:238-270): the
    samples go to the device in the file's own integer type and are converted there."""
    x, sr, ch = parse_wav(wav)
    meta = reflo_metadata(x.size, sr, ch, level, encoding_time, source_format, original_filename, tags)
    enc = Encoder(sr, ch, 16, context=context).with_compression(level)
    return enc.encode(x, meta) if x.dtype == np.float32 else enc.encode_pcm(x, meta)
