"""Byte parity at the full sizes of the BASELINE configs and on the reference's remaining edge cases.

* config 2 at its full 3600 frames: container invariants over the whole hour + 36 frames spread over the
  stream compared byte for byte with the oracle (frames are independent in the reference, encoder.rs:53-61,
  so frame g of the file equals the only frame of the oracle's encoding of second g);
* config 3 as one batch of 20 x 180 s tracks: two whole tracks compared with the oracle, CRC on all;
* 192 kHz mono + stereo and the non-standard 12 345 Hz rate (libflo/tests/rust/edge_case_tests.rs:157-169);
* the 10-minute mono stream (edge_case_tests.rs:466-473) through the host entry.
Everything goes through the C ABI (device entry for the large inputs: they are synthesised in HBM)."""
import os
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from helpers import oracle, pcm16_to_f32, synth_pcm16

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import flo_b200
    return flo_b200


@pytest.fixture(scope="module")
def ctx(fb):
    return fb.default_context(0)


def device_encode(ctx, pcm_tracks, sr, ch, bits, level):
    """interleaved int16 tensors in HBM -> list of .flo images (numpy uint8), f32 entry"""
    import torch
    xs = [p.to(torch.float32) * (1.0 / 32768.0) for p in pcm_tracks]
    n = [int(t.numel()) for t in xs]
    bound = ctx.output_bound(n, [sr] * len(n), [ch] * len(n))
    out = torch.empty(bound, dtype=torch.uint8, device=xs[0].device)
    off, ln = ctx.encode_batch_device([t.data_ptr() for t in xs], n, [sr] * len(n), [ch] * len(n), [bits] * len(n),
                                      out.data_ptr(), bound, level=level)
    host = out.cpu().numpy()
    return [host[int(o):int(o) + int(l)] for o, l in zip(off, ln)]


def check_container(img, sr, n_frames, total_samples):
    f = oracle.FloFile(img.tobytes())
    assert f.num_frames == n_frames and f.total_samples == total_samples
    assert zlib.crc32(f.data_chunk()) == f.crc32
    off = 0
    for i, fr in enumerate(f.frames):
        assert fr.byte_offset == off and fr.timestamp_ms == 1000 * i
        off += fr.frame_size
    assert off == f.data_size
    return f


def test_config2_full_hour_frames_against_oracle(ctx):
    import torch
    import synth_torch
    sr, ch, secs = 44100, 2, 3600
    pcm = synth_torch.synth_pcm16_long(secs * sr, ch, sr, 0xF10 + 2, "multitone", 64, torch.device("cuda", 0))
    img = device_encode(ctx, [pcm], sr, ch, 16, 5)[0]
    f = check_container(img, sr, secs, secs * sr)
    picks = list(range(0, secs, 100))                 # 36 frames, first to last hundred
    segs = [pcm[g * sr * ch:(g + 1) * sr * ch].cpu().numpy() for g in picks]
    with ThreadPoolExecutor(8) as ex:
        refs = list(ex.map(lambda s: oracle.encode_pcm16(s, sr, ch, 16, 5, b""), segs))
    for g, ref in zip(picks, refs):
        assert f.frame_bytes(g) == oracle.FloFile(ref).frame_bytes(0), f"frame {g} of the hour differs from the oracle"


def test_config3_batch_of_20_tracks_against_oracle(ctx):
    import torch
    import synth_torch
    sr, ch, secs, ntr = 44100, 2, 180, 20
    dev = torch.device("cuda", 0)
    pcms = [synth_torch.synth_pcm16_long(secs * sr, ch, sr, 0xF10 + 2 + 131 * t, "multitone", 64, dev) for t in range(ntr)]
    imgs = device_encode(ctx, pcms, sr, ch, 16, 5)
    assert len(imgs) == ntr
    for img in imgs:
        check_container(img, sr, secs, secs * sr)
    picks = (3, 17)
    with ThreadPoolExecutor(2) as ex:
        refs = list(ex.map(lambda t: oracle.encode_pcm16(pcms[t].cpu().numpy(), sr, ch, 16, 5, b""), picks))
    for t, ref in zip(picks, refs):
        assert imgs[t].tobytes() == ref, f"track {t} of the batch differs from the oracle"


@pytest.mark.parametrize("sr,ch,level", [(192000, 1, 5), (192000, 2, 5), (192000, 2, 9), (12345, 1, 5), (12345, 2, 9)])
def test_high_and_non_standard_sample_rates(fb, sr, ch, level):
    n = sr + sr // 3 + 7                              # a full frame, a partial one and an odd tail
    pcm = synth_pcm16(n, ch, sr, seed=0xF17 + sr + ch, kind="sweep" if sr > 100000 else "multitone", noise_lsb=24)
    x = pcm16_to_f32(pcm)
    want = oracle.encode(x, sr, ch, 16, level, b"rate")
    got = fb.Encoder(sr, ch, 16).with_compression(level).encode(x, b"rate")
    assert got == want, f"{sr} Hz x {ch}: differs from the oracle"
    assert fb.Encoder(sr, ch, 16).with_compression(level).encode_pcm16(pcm, b"rate") == want


def test_reference_sine_vectors_192k_and_12345(fb):
    """the literal inputs of edge_case_tests.rs:157-169: sin(i * 0.01) as f32, mono, 16 bits"""
    for sr in (192000, 12345):
        x = np.sin(np.arange(sr, dtype=np.float32) * np.float32(0.01)).astype(np.float32)
        want = oracle.encode(x, sr, 1, 16, 5, b"")
        got = fb.Encoder(sr, 1, 16).encode(x, b"")
        assert got == want
        dec = fb.Decoder().decode(got)
        assert dec.size == x.size and np.max(np.abs(dec - x)) <= 1.0 / 32767 + 1e-6


def test_ten_minute_mono_stream(fb):
    """edge_case_tests.rs:466-473: 10 minutes of sin(i * 0.001), mono 44.1 kHz, through the host entry"""
    n = 44100 * 600
    x = np.sin(np.arange(n, dtype=np.float32) * np.float32(0.001)).astype(np.float32)
    got = fb.Encoder(44100, 1, 16).encode(x, b"")
    f = oracle.FloFile(got)
    assert f.num_frames == 600 and f.total_samples == n and zlib.crc32(f.data_chunk()) == f.crc32
    for g in (0, 299, 599):
        ref = oracle.FloFile(oracle.encode(x[g * 44100:(g + 1) * 44100], 44100, 1, 16, 5, b""))
        assert f.frame_bytes(g) == ref.frame_bytes(0), f"frame {g}"
    dec = fb.Decoder().decode(got)
    assert dec.size == n


def test_device_entry_orders_after_async_producers(ctx):
    """The device entry must read inputs that were produced asynchronously right before the call: on torch's
    default stream (the context's own blocking stream orders after it) and on a side stream handed over with
    set_stream (include/flo_b200.h, stream ordering)."""
    import torch
    sr, ch = 44100, 2
    pcm = synth_pcm16(4 * sr, ch, sr, seed=0xF18)
    want = oracle.encode(pcm16_to_f32(pcm), sr, ch, 16, 5, b"")
    dev = torch.device("cuda", 0)
    src = torch.from_numpy(pcm.astype(np.int16)).to(dev)
    n = src.numel()
    bound = ctx.output_bound([n], [sr], [ch])
    out = torch.empty(bound, dtype=torch.uint8, device=dev)
    junk = torch.randn(64 << 20, device=dev)

    def run():
        off, ln = ctx.encode_batch_device([x.data_ptr()], [n], [sr], [ch], [16], out.data_ptr(), bound, level=5)
        return out[int(off[0]):int(off[0]) + int(ln[0])].cpu().numpy().tobytes()

    ctx.set_stream(0)
    x = torch.zeros(n, dtype=torch.float32, device=dev)
    for _ in range(4):
        junk = junk * 1.0001 + 0.5                     # keeps the default stream busy in front of the producer
    x.copy_(src.to(torch.float32) * (1.0 / 32768.0))   # asynchronous producer on the default stream
    assert run() == want
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(4):
            junk = junk * 1.0001 + 0.5
        x.zero_()
        x.copy_(src.to(torch.float32) * (1.0 / 32768.0))
        ctx.set_stream(side.cuda_stream)
        got = run()
    ctx.set_stream(0)
    side.synchronize()
    assert got == want
