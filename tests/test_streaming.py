"""Row N3: StreamingEncoder semantics (libflo/src/streaming/encoder.rs) on top of the batch encoder.
The checker is the oracle's line-by-line restatement (`oracle.StreamingEncoderRef`)."""
import zlib

import numpy as np
import pytest

from helpers import oracle, pcm16_to_f32, synth_pcm16
from flo_b200.streaming import reserialize_frame

SR = 8000


@pytest.mark.parametrize("channels,kind,level", [(1, "multitone", 5), (2, "speech", 9), (2, "multitone", 0), (3, "sweep", 3),
                                                 (2, "noise", 5), (2, "silence", 5), (1, "short", 5)])
def test_reserialize_matches_restatement_cpu(channels, kind, level):
    """Host logic only: the product's frame re-serialiser against the restatement, on oracle-encoded frames."""
    if kind == "noise":
        x = pcm16_to_f32(np.random.default_rng(1).integers(-32768, 32768, SR * channels).astype(np.int16))
    elif kind == "silence":
        x = np.zeros(SR * channels, np.float32)
    elif kind == "short":
        x = pcm16_to_f32(synth_pcm16(777, channels, SR, seed=2))
    else:
        x = pcm16_to_f32(synth_pcm16(SR, channels, SR, seed=3, kind=kind))
    ref = oracle.StreamingEncoderRef(SR, channels, 16).with_compression(level)
    assert reserialize_frame(oracle.encode(x, SR, channels, 16, level, b""), channels) == ref._encode_frame_data(x)


def test_container_crc_is_zlib_crc32():
    d = bytes(range(256)) * 37
    assert oracle.crc32(d) == zlib.crc32(d) & 0xFFFFFFFF


def _same_frame(a, b):
    return a is not None and b is not None and (a.index, a.timestamp_ms, a.samples, a.data) == \
        (b["index"], b["timestamp_ms"], b["samples"], b["data"])


@pytest.mark.gpu
def test_reference_scenario_two_and_a_half_seconds():
    """libflo/src/streaming/tests.rs:55-82."""
    import flo_b200
    x = np.sin(np.arange(SR * 5 // 2, dtype=np.float32) * np.float32(0.01)).astype(np.float32)
    enc, ref = flo_b200.StreamingEncoder(SR, 1, 16), oracle.StreamingEncoderRef(SR, 1, 16)
    enc.push_samples(x); ref.push_samples(x)
    assert enc.pending_frames() == 2 == ref.pending_frames() and enc.pending_samples() == SR // 2
    assert _same_frame(enc.next_frame(), ref.next_frame()) and _same_frame(enc.next_frame(), ref.next_frame())
    assert enc.next_frame() is None
    out = enc.finalize(b"")
    assert out == ref.finalize(b"") and out[4:6] == bytes([1, 2]) and enc.pending_frames() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("channels,level,chunk", [(1, 5, 1234), (2, 5, 16000), (2, 9, 777), (2, 2, 40001), (6, 4, 5000)])
def test_chunked_pushes_match(channels, level, chunk):
    import flo_b200
    x = pcm16_to_f32(synth_pcm16(int(SR * 4.3), channels, SR, seed=channels + level, kind="speech"))
    enc = flo_b200.StreamingEncoder(SR, channels, 16).with_compression(level)
    ref = oracle.StreamingEncoderRef(SR, channels, 16).with_compression(level)
    taken = 0
    for s in range(0, x.size, chunk):
        enc.push_samples(x[s:s + chunk]); ref.push_samples(x[s:s + chunk])
        assert enc.pending_frames() == ref.pending_frames() and enc.pending_samples() == ref.pending_samples()
        if s // chunk % 3 == 1:                                    # take a frame out now and then
            a, b = enc.next_frame(), ref.next_frame()
            assert (a is None) == (b is None)
            if a is not None:
                assert _same_frame(a, b); taken += 1
    assert taken > 0 or chunk > x.size // 3
    fa, fb = enc.flush(), ref.flush()
    assert _same_frame(fa, fb) and fa.samples == (x.size // channels) % SR
    assert enc.flush() is None and ref.flush() is None
    meta = b"\x81\xa5title\xa1x"
    assert enc.finalize(meta) == ref.finalize(meta)


@pytest.mark.gpu
def test_many_frames_in_one_push_use_one_device_pass():
    import flo_b200
    x = pcm16_to_f32(synth_pcm16(SR * 30, 2, SR, seed=8))
    enc, ref = flo_b200.StreamingEncoder(SR, 2, 16), oracle.StreamingEncoderRef(SR, 2, 16)
    enc.push_samples(x); ref.push_samples(x)
    assert enc.pending_frames() == 30
    assert flo_b200.default_context().last_timing()["launches"] <= 6          # one batch call, not 30
    assert enc.finalize(b"meta") == ref.finalize(b"meta")
