"""Row N2: the GPU lossless decoder against the oracle's restatement of Reader::read +
Decoder::decode (libflo/src/reader.rs, libflo/src/lossless/decoder.rs).  Bit-exact f32 output; same error
messages.  Files come from three places: the reference's shipped examples, the GPU encoder, and hand-built
images that reach the decoder arms the encoder never produces."""
import struct

import numpy as np
import pytest

from helpers import LOSSLESS_EXAMPLES, golden_bytes, oracle, pcm16_to_f32, synth_pcm16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import flo_b200
    return flo_b200


def same_f32(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def check(fb, data: bytes):
    want = oracle.decode(data)
    got = fb.Decoder().decode(data)
    assert got.dtype == np.float32 and same_f32(np.asarray(got), want)
    return want


# ---- hand-built file images -------------------------------------------------------------------------------------
def rice_channel(res, k, order=0, coeffs=(), shift=128, enc=0, n_coeffs_written=None):
    """One ALPC channel body: order, coeffs, shift, encoding byte, [k], Rice bits (writer.rs layout)."""
    body = bytes([order])
    for c in list(coeffs)[: (len(coeffs) if n_coeffs_written is None else n_coeffs_written)]:
        body += struct.pack("<i", c)
    body += bytes([shift, enc])
    if enc == 0:
        body += bytes([k])
    return body + oracle.rice_encode_i32(np.asarray(res, np.int32), k)


def build_file(channels, frames, sample_rate=8000, meta=b"", data_size=None, toc_offsets=None):
    """frames: [(type, n, flags, [channel payload bytes ...])] -> file image (header 70 B, TOC, DATA, META)."""
    blobs = []
    for t, n, flags, chans in frames:
        b = bytes([t]) + struct.pack("<I", n) + bytes([flags])
        for ch in chans:
            b += struct.pack("<I", len(ch)) + ch
        blobs.append(b)
    toc = struct.pack("<I", len(blobs))
    off = 0
    for i, b in enumerate(blobs):
        o = off if toc_offsets is None else toc_offsets[i]
        toc += struct.pack("<IQII", i, o, len(b), 0)
        off += len(b)
    data = b"".join(blobs)
    total = sum(f[1] for f in frames)
    hdr = b"FLO!" + bytes([1, 1]) + struct.pack("<H", 0) + struct.pack("<I", sample_rate) + bytes([channels, 16])
    hdr += struct.pack("<Q", total) + bytes([5, 0, 0, 0]) + struct.pack("<I", 0)
    hdr += struct.pack("<QQQQQ", 66, len(toc), len(data) if data_size is None else data_size, 0, len(meta))
    assert len(hdr) == 70
    return hdr + toc + data + meta


# ---- reference files and encoder output ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", LOSSLESS_EXAMPLES)
def test_shipped_examples(fb, name):
    data = golden_bytes(name)
    want = check(fb, data)
    samples, info = fb.default_context().decode(data)
    f = oracle.FloFile(data)
    assert info["sample_rate"] == f.sample_rate and info["channels"] == f.channels and info["n_frames"] == f.num_frames
    assert info["total_samples"] == f.total_samples and info["meta_size"] == f.meta_size and info["level"] == f.level
    assert data[info["meta_offset"]:info["meta_offset"] + info["meta_size"]] == data[len(data) - f.meta_size:]
    assert samples.size == want.size


@pytest.mark.parametrize("channels,level,kind,seconds", [
    (1, 5, "multitone", 2.5), (2, 5, "multitone", 3.2), (2, 9, "speech", 2.0), (2, 0, "sweep", 1.5), (1, 2, "speech", 1.1),
    (6, 5, "multitone", 1.3), (3, 7, "sweep", 2.0), (2, 3, "tone", 1.0), (8, 4, "speech", 1.2),
])
def test_encoder_output(fb, channels, level, kind, seconds):
    sr = 16000
    pcm = synth_pcm16(int(sr * seconds), channels, sr, seed=channels * 131 + level, kind=kind)
    data = fb.Encoder(sr, channels, 16).with_compression(level).encode_pcm16(pcm)
    want = check(fb, data)
    if level > 0:       # level 0 writes frame type 0 (= Silence) for its frames (encoder.rs frame-type rule): the reference loses them too
        q = np.trunc(np.clip(pcm16_to_f32(pcm) * np.float32(32767.0), -32767, 32767)).astype(np.int32)
        assert same_f32(want, q.astype(np.float32) * np.float32(1.0 / 32767.0))    # lossless round trip


def test_noise_silence_and_spikes(fb):
    sr = 8000
    rng = np.random.default_rng(11)
    noise = rng.integers(-32768, 32768, sr * 2, dtype=np.int64).astype(np.int16)       # raw-coded frames
    silence = np.zeros(sr * 2, np.int16)                                                 # silence frames
    spikes = np.zeros(sr * 2, np.int16); spikes[::997] = 32767; spikes[5::1499] = -32768  # long unary codes
    pcm = np.concatenate([noise, silence, spikes, noise[:1234]])
    for ch in (1, 2):
        data = fb.Encoder(sr, ch, 16).encode_pcm16(pcm[: pcm.size // ch * ch])
        check(fb, data)


def test_full_size_round_trip(fb):
    """120 s of 44.1 kHz stereo: decode(encode(x)) is the quantised input (size-independent property)."""
    sr, ch = 44100, 2
    pcm = synth_pcm16(sr * 120, ch, sr, seed=5)
    data = fb.Encoder(sr, ch, 16).encode_pcm16(pcm)
    got = fb.Decoder().decode(data)
    q = np.trunc(np.clip(pcm16_to_f32(pcm) * np.float32(32767.0), -32767, 32767)).astype(np.int32)
    assert same_f32(np.asarray(got), q.astype(np.float32) * np.float32(1.0 / 32767.0))
    t = fb.default_context().last_timing()
    assert t["launches"] == 3


# ---- decoder arms the encoder never takes -------------------------------------------------------------------------
def test_fixed_orders_wrapping(fb):
    rng = np.random.default_rng(3)
    n = 700
    frames = []
    for order in (0, 1, 2, 3, 4, 5, 9, 127):
        res = rng.integers(-3000, 3000, n)
        res[:6] = rng.integers(-(1 << 20), 1 << 20, 6)
        frames.append((4, n, 0, [rice_channel(res, 7, shift=128 + order)]))
    big = rng.integers(-(1 << 27), 1 << 27, n)                                            # sums wrap around i32
    frames.append((4, n, 0, [rice_channel(big, 24, shift=128 + 4)]))
    check(fb, build_file(1, frames))


def test_lpc_arbitrary_coefficients_and_shifts(fb):
    rng = np.random.default_rng(4)
    frames = []
    for order in range(1, 13):
        n = 300 + 17 * order
        coeffs = rng.integers(-20000, 20000, order).tolist()
        res = rng.integers(-200, 200, n)
        for shift in (15, 12, 0, 31, 40, 63, 64 + 15, 127):
            frames.append((order, n, 0, [rice_channel(res, 5, order=order, coeffs=coeffs, shift=shift)]))
    wild = rng.integers(-(1 << 31), 1 << 31, 12).tolist()                                 # i64 accumulators wrap
    frames.append((12, 500, 0, [rice_channel(rng.integers(-(1 << 24), 1 << 24, 500), 22, order=12, coeffs=wild, shift=3)]))
    check(fb, build_file(1, frames))


def test_mixed_channels_in_one_warp(fb):
    """Every lane of a warp in a different mode / order / length."""
    rng = np.random.default_rng(5)
    frames = []
    for i in range(40):
        n = int(rng.integers(0, 900))
        chans = []
        for c in range(3):
            sel = int(rng.integers(0, 6))
            res = rng.integers(-500, 500, n)
            if sel == 0:
                chans.append(rice_channel(res, int(rng.integers(0, 12)), shift=128 + int(rng.integers(0, 5))))
            elif sel == 1:
                o = int(rng.integers(1, 13))
                chans.append(rice_channel(res, int(rng.integers(0, 12)), order=o, coeffs=rng.integers(-9000, 9000, o).tolist(), shift=15))
            elif sel == 2:
                chans.append(bytes([0, 3, 2]) + rng.integers(0, 256, int(rng.integers(0, 2 * n + 9)), dtype=np.uint8).tobytes())  # PCM in ALPC
            elif sel == 3:
                chans.append(bytes([0]))                                                  # header reads run into the next channel
            elif sel == 4:
                chans.append(rice_channel(res, 0, shift=130, enc=1))                        # Golomb byte: no k, k = 0
            else:
                chans.append(rice_channel(res, 9, order=6, coeffs=rng.integers(-9000, 9000, 6).tolist(), shift=14, n_coeffs_written=6)[: 1 + 4 * int(rng.integers(0, 6))])
        frames.append((int(rng.integers(1, 13)), n, int(rng.integers(0, 2)), chans))
    data = build_file(3, frames) + b"\0" * 16
    check(fb, data)


def test_truncated_residuals_and_unary_cap(fb):
    n = 400
    res = np.arange(n) % 37 - 18
    full = rice_channel(res, 3, shift=129)
    frames = [(2, n, 0, [full[: len(full) // 2]]),                                        # bits run out: zeros (rice.rs:128-131)
              (2, n, 0, [bytes([0, 129, 0, 2]) + b"\xff" * 40 + b"\x00\x12\x34"]),        # 320 ones: quotient capped at 256 reads
              (2, n, 0, [bytes([0, 129, 0, 0]) + b"\xff" * 31 + b"\xfe" + b"\xaa" * 8]),  # exactly 255 ones then the terminator
              (2, n, 0, [bytes([0, 128, 0, 31]) + bytes(range(1, 200))]),                # k = 31
              (2, n, 0, [bytes([0, 128, 0, 0]) + b"\xff" * 7])]                          # ends inside a unary run
    check(fb, build_file(1, frames))


def test_raw_silence_reserved_and_mid_side(fb):
    rng = np.random.default_rng(6)
    n = 333
    raw_l = rng.integers(-32768, 32768, n).astype("<i2").tobytes()
    raw_short = raw_l[:101]                                                               # odd byte count, padded with zeros
    m = rng.integers(-60000, 60000, n); s = rng.integers(-60000, 60000, n)                # odd sums: `/ 2` truncates toward zero
    frames = [(254, n, 0, [raw_l, raw_short]),
              (254, n, 1, [raw_l, raw_short]),
              (0, n, 1, [b"", b"junk"]),
              (200, n, 0, [b"abc", b""]),                                                 # reserved type: silence
              (5, n, 1, [rice_channel(m, 12, shift=128), rice_channel(s, 12, shift=128)]),
              (5, 0, 1, [rice_channel([], 0, shift=128), rice_channel([], 0, shift=128)]),
              (5, 1, 1, [rice_channel([-7], 2, shift=131), rice_channel([4], 2, shift=131)])]
    check(fb, build_file(2, frames))


def test_toc_break_and_overlapping_offsets(fb):
    n = 64
    ch = rice_channel(np.arange(n) - 32, 4, shift=129)
    frames = [(1, n, 0, [ch])] * 5
    data = build_file(1, frames, toc_offsets=[0, 2 * (10 + len(ch)), 0, 10 ** 9, 10 + len(ch)])   # 4th entry stops the reader
    want = check(fb, data)
    assert want.size == 3 * n
    assert check(fb, build_file(1, frames, data_size=10 + len(ch) + 3)).size == 2 * n     # frame starting before data_end is read
    bad = (3, n, 0, [bytes([13]) + ch[1:]])                                               # invalid LPC order ...
    flen = 10 + len(ch)
    assert check(fb, build_file(1, [frames[0], frames[0], bad], toc_offsets=[0, 10 ** 9, 2 * flen])).size == n   # ... behind the stop entry: never read


def test_empty_and_degenerate_files(fb):
    assert check(fb, build_file(1, [])).size == 0
    assert check(fb, build_file(0, [(1, 10, 0, [])])).size == 0
    hdr_only = build_file(1, [])[:70]
    hdr_only = hdr_only[:38] + struct.pack("<Q", 0) + hdr_only[46:]                        # toc_size 0: no TOC at all
    assert check(fb, hdr_only).size == 0
    assert check(fb, build_file(2, [(0, 0, 0, [b"", b""])], meta=b"tail")).size == 0


@pytest.mark.parametrize("mutate,msg", [
    (lambda d: b"FLO?" + d[4:], "Invalid flo file: bad magic"),
    (lambda d: d[:3], "Invalid flo file: bad magic"),
    (lambda d: d[:60], "Unexpected end of file"),
    (lambda d: d[:72], "Unexpected end of file"),
    (lambda d: d[:70] + struct.pack("<I", 100001) + d[74:], "Invalid TOC: too many entries"),
    (lambda d: d[:70] + struct.pack("<I", 50000) + d[74:], "Unexpected end of file"),
    (lambda d: d[: len(d) - 40], "Unexpected end of file"),
    (lambda d: d[:62] + struct.pack("<Q", 9) + d[70:], "Unexpected end of file"),          # META longer than the file
])
def test_errors_match_the_reader(fb, mutate, msg):
    sr = 8000
    good = fb.Encoder(sr, 2, 16).encode_pcm16(synth_pcm16(sr * 3, 2, sr, seed=9))
    bad = mutate(good)
    with pytest.raises(ValueError) as eo:
        oracle.decode(bad)
    assert str(eo.value) == msg
    with pytest.raises(fb.FloError) as eg:
        fb.Decoder().decode(bad)
    assert str(eg.value) == msg


def test_frame_level_errors(fb):
    n = 50
    ch = rice_channel(np.arange(n), 3, shift=129)
    cases = [(build_file(1, [(1, n, 0, [ch]), (3, n, 0, [bytes([13]) + ch[1:]])]), "Invalid LPC order"),
             (build_file(1, [(1, 2000001, 0, [ch])]), "Invalid frame: too many samples"),
             (build_file(1, [(1, n, 0, [ch]), (1, n, 0, [ch])])[:-3], "Unexpected end of file"),
             (build_file(2, [(1, n, 0, [ch])]), "Unexpected end of file"),                 # second channel missing
             # two errors in one file: the reader meets the bad order of frame 1 before the truncated frame 2 ...
             (build_file(1, [(1, n, 0, [ch]), (3, n, 0, [bytes([13]) + ch[1:]]), (1, n, 0, [ch])])[:-3], "Invalid LPC order"),
             # ... and inside one channel the order byte is tested before its payload is found to run off the file
             (build_file(1, [(1, n, 0, [ch]), (3, n, 0, [bytes([13]) + ch[1:]])])[:-5], "Invalid LPC order"),
             # the other way round: a truncated frame 1 in front of a bad order in frame 2 cannot be built (frames are
             # contiguous), but a frame that claims too many samples in front of a bad order reports the former
             (build_file(1, [(1, 2000001, 0, [ch]), (3, n, 0, [bytes([13]) + ch[1:]])]), "Invalid frame: too many samples")]
    for data, msg in cases:
        with pytest.raises(ValueError) as eo:
            oracle.decode(data)
        assert str(eo.value) == msg
        with pytest.raises(fb.FloError) as eg:
            fb.Decoder().decode(data)
        assert str(eg.value) == msg
    with pytest.raises(fb.FloError):
        fb.Decoder().decode(build_file(1, [(253, n, 0, [b"blob"])]))                      # transform frames: refused


def test_device_resident_round_trip(fb):
    """encode_batch_device -> flo_decode_device without the samples leaving the GPU."""
    import torch
    sr, ch = 44100, 2
    pcm = synth_pcm16(sr * 20, ch, sr, seed=21)
    ctx = fb.default_context()
    d_in = torch.from_numpy(pcm).cuda()
    bound = ctx.output_bound([pcm.size], [sr], [ch])
    d_file = torch.empty(bound, dtype=torch.uint8, device="cuda")
    offs, lens = ctx.encode_batch_device([d_in.data_ptr()], [pcm.size], [sr], [ch], [16], d_file.data_ptr(), bound, level=5, fmt=fb.FMT_PCM16)
    d_out = torch.empty(pcm.size, dtype=torch.float32, device="cuda")
    n, info = ctx.decode_device(d_file.data_ptr() + int(offs[0]), int(lens[0]), d_out.data_ptr(), d_out.numel())
    assert n == pcm.size and info["channels"] == ch and info["decoded_frames"] == pcm.size // ch
    q = torch.trunc(torch.clamp(d_in.float() * (1.0 / 32768.0) * 32767.0, -32767, 32767))
    assert torch.equal(d_out, q * torch.tensor(1.0 / 32767.0, dtype=torch.float32, device="cuda"))
    with pytest.raises(fb.FloError):
        ctx.decode_device(d_file.data_ptr() + int(offs[0]), int(lens[0]), d_out.data_ptr(), 10)


def test_decode_on_caller_stream_and_concurrent_contexts(fb):
    """Two contexts decoding from two threads, one of them on a caller-owned stream; results stay bit-exact."""
    import threading
    import torch
    sr = 22050
    files = [fb.Encoder(sr, 2, 16).with_compression(lv).encode_pcm16(synth_pcm16(sr * 6 + 11 * lv, 2, sr, seed=40 + lv)) for lv in (3, 5, 8)]
    wants = [oracle.decode(f) for f in files]
    ctxs = [fb.Context(0), fb.Context(0)]
    stream = torch.cuda.Stream()
    ctxs[1].set_stream(stream.cuda_stream)
    errs = []

    def work(ci):
        try:
            for rep in range(6):
                for f, w in zip(files, wants):
                    got, info = ctxs[ci].decode(f)
                    assert same_f32(np.asarray(got), w) and info["channels"] == 2
        except Exception as e:          # surfaced in the main thread
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    ctxs[1].set_stream(0)
    for c in ctxs:
        c.close()
    assert not errs, errs


def test_i16_output_entry(fb):
    """flo_decode_i16: the integer samples before i32_to_f32 (decoder.rs:61-72), saturated to i16 -- against the oracle's
    integer decode; and its relation to the f32 entry."""
    import torch
    for sr, ch, level, n in ((44100, 2, 5, 3 * 44100 + 17), (8000, 1, 9, 20001), (48000, 6, 3, 30000), (44100, 2, 0, 50000)):
        pcm = synth_pcm16(n, ch, sr, seed=0xD16 + ch + level, kind="multitone", noise_lsb=64)
        img = fb.Encoder(sr, ch, 16).with_compression(level).encode_pcm16(pcm, b"")
        want = np.clip(oracle.decode_i32(img), -32768, 32767).astype(np.int16)
        got = fb.Decoder().decode_i16(img)
        assert got.dtype == np.int16 and np.array_equal(got, want)
        f32 = fb.Decoder().decode(img)
        ok = np.abs(want.astype(np.int32)) <= 32767
        assert np.array_equal(f32[ok], (want[ok].astype(np.float32) * np.float32(1.0 / 32767.0)))
    # saturation: a hand-built file whose samples leave the 16-bit range
    big = rice_channel(np.array([40000, -40000, 123, -32768, 32767, 70000], np.int64), 12, shift=128)
    data = build_file(1, [(1, 6, 0, [big])])
    assert fb.Decoder().decode_i16(data).tolist() == [32767, -32768, 123, -32768, 32767, 32767]
    # device-resident entry and the capacity error
    ctx = fb.default_context(0)
    dev = torch.device("cuda", 0)
    d_file = torch.from_numpy(np.frombuffer(img, np.uint8).copy()).to(dev)
    d_out = torch.empty(want.size, dtype=torch.int16, device=dev)
    n_dec, info = ctx.decode_i16_device(d_file.data_ptr(), len(img), d_out.data_ptr(), d_out.numel())
    assert n_dec == want.size and np.array_equal(d_out.cpu().numpy(), want) and info["channels"] == 2
    with pytest.raises(fb.FloError):
        ctx.decode_i16_device(d_file.data_ptr(), len(img), d_out.data_ptr(), want.size - 1)
    with pytest.raises(fb.FloError) as e:
        fb.Decoder().decode_i16(img[: len(img) - 40])
    assert str(e.value) == "Unexpected end of file"
