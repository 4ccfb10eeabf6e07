"""Shared helpers for the test-suite (oracle access, golden fixtures, signals)."""
from __future__ import annotations

import gzip
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import flo_oracle as oracle  # noqa: E402  (test infrastructure: the checker)

GOLDEN = os.path.join(ROOT, "tests", "golden", "examples")
LOSSLESS_EXAMPLES = [
    "audio_lossless.flo", "chord_cmajor_stereo.flo", "click_track_120bpm.flo", "dtmf_tones.flo",
    "hires_96khz.flo", "multitone_stereo.flo", "silence_1sec.flo", "sine_440hz_mono.flo",
    "sweep_20_20k.flo", "telephone_8khz.flo", "white_noise.flo",
]


def golden_bytes(name: str) -> bytes:
    return open(os.path.join(GOLDEN, name), "rb").read()


def golden_audio_wav_f32() -> tuple[np.ndarray, int, int]:
    """Examples/audio.wav: IEEE-float stereo 44.1 kHz; returns (interleaved f32, sr, channels)."""
    raw = gzip.open(os.path.join(GOLDEN, "audio.wav.gz"), "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, sz = raw[pos:pos + 4], struct.unpack_from("<I", raw, pos + 4)[0]
        body = raw[pos + 8:pos + 8 + sz]
        if cid == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", body, 0)
        elif cid == b"data":
            data = body
        pos += 8 + sz + (sz & 1)
    tag, ch, sr, _, _, bits = fmt
    assert tag == 3 and bits == 32, (tag, bits)
    return np.frombuffer(data, dtype="<f4").copy(), sr, ch


def coded_channels_of_frame(f: "oracle.FloFile", i: int):
    """Coded-domain int32 channels of frame i of a reference file.

    Handles the reference quirk that a Raw-typed (254) frame may hold fixed-0 Rice bytes
    (encoder.rs:115-119 + writer.rs:267-270): there the channel size is not 2*n and the
    Rice parameter is lost, so it is recovered by the unique k that round-trips."""
    fr = f.frames[i]
    coded = f.decode_frame_coded(i)
    chans = [coded[c] for c in range(len(fr.channels))]
    if fr.frame_type != 254:
        return chans
    orig = f.frame_bytes(i)
    pos = 6
    for c, ci in enumerate(fr.channels):
        sz = int.from_bytes(orig[pos:pos + 4], "little")
        body = orig[pos + 4:pos + 4 + sz]
        pos += 4 + sz
        if sz != 2 * fr.frame_samples:
            found = None
            for k in range(16):
                r = oracle.rice_decode_i32(body, k, fr.frame_samples)
                if oracle.rice_encode_i32(r, k) == body and oracle.estimate_rice_parameter_i32(r) == k:
                    found = r
                    break
            assert found is not None, "raw-typed frame with undecodable Rice payload"
            chans[c] = found
    return chans


def f32_for_ints(v: np.ndarray, silent: bool = False) -> np.ndarray:
    """f32 samples x with f32_to_i32(x) == v (x = (v +- 0.5)/32767).

    Zero maps to 1e-6 (non-silent but quantises to 0) unless the frame must be silent."""
    v = np.asarray(v, dtype=np.int64)
    x = (v + 0.5 * np.sign(v)) / 32767.0
    x = x.astype(np.float32)
    if not silent:
        x[v == 0] = np.float32(1e-6)
    return x


def file_to_f32_input(f: "oracle.FloFile") -> np.ndarray:
    """Interleaved f32 input that makes Encoder::encode reproduce reference file f
    (only valid for files without mid/side frames; the shipped examples have none)."""
    parts = []
    for i, fr in enumerate(f.frames):
        assert fr.flags == 0
        chans = coded_channels_of_frame(f, i)
        il = np.zeros(fr.frame_samples * f.channels, np.float32)
        for c in range(f.channels):
            il[c::f.channels] = f32_for_ints(chans[c], silent=(fr.frame_type == 0))
        parts.append(il)
    return np.concatenate(parts) if parts else np.zeros(0, np.float32)


# ---- deterministic integer-only test signals (identical on host and device) ----
def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_pcm16(n: int, channels: int, sample_rate: int, seed: int = 0xF10, kind: str = "multitone",
                noise_lsb: int = 64) -> np.ndarray:
    """Interleaved int16 PCM [n*channels]: multitone + hash noise, integer phase accumulators."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        table = np.round(np.sin(np.arange(4096) * (2 * np.pi / 4096)) * 16384).astype(np.int64)
        out = np.zeros((n, channels), np.int64)
        for c in range(channels):
            acc = np.zeros(n, np.int64)
            if kind == "multitone":
                freqs = [(220 + 37 * c, 3), (1330 + 101 * c, 4), (5170 + 13 * c, 5), (97, 3)]
            elif kind == "sweep":
                freqs = []
                ph = ((idx * idx * np.uint64(max(1, (1 << 32) // max(1, 4 * n)))) >> np.uint64(20)) & np.uint64(4095)
                acc += table[ph.astype(np.int64)] // 2
            elif kind == "speech":
                freqs = [(140 + 11 * c, 2), (710, 3), (1220, 4)]
            else:
                freqs = [(440, 2)]
            for f, sh in freqs:
                step = np.uint64((f << 32) // sample_rate)
                ph = ((idx * step) >> np.uint64(20)) & np.uint64(4095)
                acc += table[ph.astype(np.int64)] >> sh
            if kind == "speech":
                env = (table[((idx * np.uint64((4 << 32) // sample_rate)) >> np.uint64(20)).astype(np.int64) & 4095] + 16384) >> 7
                acc = (acc * env) >> 8
            h = _splitmix64(idx ^ np.uint64((seed + 7919 * c) << 40 & 0xFFFFFFFFFFFFFFFF))
            noise = (h & np.uint64(2 * noise_lsb - 1)).astype(np.int64) - noise_lsb
            out[:, c] = acc + noise
        if channels == 2:
            out[:, 1] = (out[:, 0] * 13) // 16 + out[:, 1] // 4
        return np.clip(out, -32768, 32767).astype(np.int16).reshape(-1)


def pcm16_to_f32(pcm: np.ndarray) -> np.ndarray:
    """reflo/src/audio.rs:247-254: s as f32 * (1/32768)."""
    return pcm.astype(np.float32) * np.float32(1.0 / 32768.0)
