"""Row N1: the reflo metadata tail.  The golden values are the META chunks of the reference's own
shipped files (written by reflo 0.1.2): rebuilding them from (length, level, time stamp, source format) must give
the same bytes."""
import numpy as np
import pytest

from helpers import LOSSLESS_EXAMPLES, file_to_f32_input, golden_audio_wav_f32, golden_bytes, oracle
from flo_b200 import reflo


def _meta_fields(meta: bytes):
    """Minimal MessagePack reader for the flat maps reflo writes (str keys, str / uint values)."""
    pos = 1
    n = meta[0] & 0x0F
    assert meta[0] & 0xF0 == 0x80

    def rd():
        nonlocal pos
        t = meta[pos]
        pos += 1
        if t < 0x80:
            return t
        if 0xA0 <= t <= 0xBF:
            s = meta[pos:pos + (t & 31)]; pos += t & 31
            return s.decode()
        if t == 0xD9:
            ln = meta[pos]; s = meta[pos + 1:pos + 1 + ln]; pos += 1 + ln
            return s.decode()
        if t in (0xCC, 0xCD, 0xCE, 0xCF):
            w = {0xCC: 1, 0xCD: 2, 0xCE: 4, 0xCF: 8}[t]
            v = int.from_bytes(meta[pos:pos + w], "big"); pos += w
            return v
        raise AssertionError(hex(t))

    out = {}
    for _ in range(n):
        k = rd(); out[k] = rd()
    assert pos == len(meta)
    return out


@pytest.mark.parametrize("name", LOSSLESS_EXAMPLES)
def test_metadata_bytes_of_shipped_examples(name):
    gold = golden_bytes(name)
    f = oracle.FloFile(gold)
    meta = gold[len(gold) - f.meta_size:]
    fields = _meta_fields(meta)
    assert fields["flo_encoder_version"] == reflo.REFLO_VERSION and fields["encoder_settings"] == "Lossless, level 5"
    rebuilt = reflo.reflo_metadata(int(f.total_samples) * f.channels, f.sample_rate, f.channels, f.level,
                                   fields["encoding_time"], fields.get("source_format"), fields.get("original_filename"))
    assert rebuilt == meta


def test_msgpack_scalars_and_order():
    m = reflo.metadata_to_msgpack({"source_format": "WAV", "title": "x" * 40, "track_number": 300, "play_count": 1 << 33,
                                   "length_ms": 70000})
    # declaration order: title, track_number, length_ms, play_count, source_format
    assert m[0] == 0x85 and m[1:7] == b"\xa5title" and m[7:9] == b"\xd9\x28"
    assert b"\xactrack_number\xcd\x01\x2c" in m and b"\xa9length_ms\xce\x00\x01\x11\x70" in m
    assert b"\xaaplay_count\xcf\x00\x00\x00\x02\x00\x00\x00\x00" in m and m.endswith(b"\xadsource_format\xa3WAV")
    assert m.index(b"track_number") < m.index(b"length_ms") < m.index(b"play_count") < m.index(b"source_format")
    with pytest.raises(Exception):
        reflo.metadata_to_msgpack({"pictures": 1})


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOSSLESS_EXAMPLES)
def test_whole_file_equals_reflo_output(name):
    """encode_from_samples with the file's own time stamp reproduces the reference file completely."""
    gold = golden_bytes(name)
    f = oracle.FloFile(gold)
    fields = _meta_fields(gold[len(gold) - f.meta_size:])
    x = file_to_f32_input(f)
    out = reflo.encode_from_samples(x, f.sample_rate, f.channels, f.level, encoding_time=fields["encoding_time"],
                                    source_format=fields.get("source_format"), original_filename=fields.get("original_filename"))
    assert out == gold


@pytest.mark.gpu
def test_config1_audio_wav_all_246_bytes():
    x, sr, ch = golden_audio_wav_f32()
    gold = golden_bytes("audio_lossless.flo")
    out = reflo.encode_from_samples(x, sr, ch, 5, encoding_time="2026-03-09T20:46:05Z", source_format="UNKNOWN")
    assert len(out) == 246 and out == gold


def _wav(samples: np.ndarray, sr: int, ch: int, tag: int) -> bytes:
    import struct
    bits = samples.dtype.itemsize * 8
    fmt = struct.pack("<HHIIHH", tag, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits)
    body = samples.tobytes()
    return b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + \
        b"LIST" + struct.pack("<I", 4) + b"INFO" + b"data" + struct.pack("<I", len(body)) + body


def test_parse_wav_types():
    for arr, tag in [(np.arange(200, dtype=np.uint8), 1), (np.arange(-100, 100, dtype=np.int16), 1),
                     (np.arange(-100, 100, dtype=np.int32) * 70000, 1), (np.linspace(-1, 1, 200).astype(np.float32), 3)]:
        x, sr, ch = reflo.parse_wav(_wav(arr, 22050, 2, tag))
        assert (sr, ch) == (22050, 2) and x.dtype == arr.dtype and np.array_equal(x, arr)
    with pytest.raises(Exception):
        reflo.parse_wav(b"RIFF0000WAVE")
    with pytest.raises(Exception):
        reflo.parse_wav(_wav(np.zeros(30, np.uint8), 8000, 1, 85))          # compressed format tag


@pytest.mark.gpu
def test_encode_wav_reproduces_the_shipped_file():
    """`reflo encode audio.wav` -> Examples/audio_lossless.flo, all 246 bytes, from the WAV bytes themselves."""
    import gzip
    import os
    from helpers import GOLDEN
    wav = gzip.open(os.path.join(GOLDEN, "audio.wav.gz"), "rb").read()
    out = reflo.encode_wav(wav, 5, encoding_time="2026-03-09T20:46:05Z", source_format="UNKNOWN")
    assert out == golden_bytes("audio_lossless.flo")


@pytest.mark.gpu
def test_encode_wav_integer_arms():
    sr = 8000
    rng = np.random.default_rng(7)
    for arr in [rng.integers(0, 256, sr * 2 + 5, dtype=np.int64).astype(np.uint8),
                (rng.integers(-2**31, 2**31, sr * 2 + 4, dtype=np.int64)).astype(np.int32) >> 4]:
        x = oracle.u8_to_f32(arr) if arr.dtype == np.uint8 else oracle.s32_to_f32(arr)
        meta = reflo.reflo_metadata(arr.size // 2 * 2, sr, 2, 5, "2026-01-01T00:00:00Z", "WAV")
        assert reflo.encode_wav(_wav(arr, sr, 2, 1), 5, encoding_time="2026-01-01T00:00:00Z") == \
            oracle.encode(x[: arr.size // 2 * 2], sr, 2, 16, 5, meta)
