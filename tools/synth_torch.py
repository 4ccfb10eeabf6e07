"""Device-side twin of tests/helpers.synth_pcm16 (integer-only, bit-identical to the numpy version).

Bench/test infrastructure: generates the synthetic PCM of BASELINE.json's configs directly in HBM so
that a 1-hour stream does not have to be built on the host.  torch is plumbing here (device memory +
elementwise integer ops), not part of the encode path."""
from __future__ import annotations

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _i64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def _splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def sine_table(device) -> torch.Tensor:
    t = np.round(np.sin(np.arange(4096) * (2 * np.pi / 4096)) * 16384).astype(np.int64)
    return torch.from_numpy(t).to(device)


def synth_pcm16(n: int, channels: int, sample_rate: int, seed: int = 0xF10, kind: str = "multitone",
                noise_lsb: int = 64, device="cuda", start: int = 0, total: int | None = None) -> torch.Tensor:
    """Interleaved int16 [n*channels] for sample-frame indices start..start+n of an n_total-long signal."""
    total = n if total is None else total
    table = sine_table(device)
    idx = torch.arange(start, start + n, dtype=torch.int64, device=device)
    cols = []
    for c in range(channels):
        acc = torch.zeros(n, dtype=torch.int64, device=device)
        if kind == "multitone":
            freqs = [(220 + 37 * c, 3), (1330 + 101 * c, 4), (5170 + 13 * c, 5), (97, 3)]
        elif kind == "sweep":
            freqs = []
            mul = max(1, (1 << 32) // max(1, 4 * total))
            ph = _lsr(idx * idx * mul, 20) & 4095
            acc += torch.div(table[ph], 2, rounding_mode="floor")
        elif kind == "speech":
            freqs = [(140 + 11 * c, 2), (710, 3), (1220, 4)]
        else:
            freqs = [(440, 2)]
        for f, sh in freqs:
            step = (f << 32) // sample_rate
            ph = _lsr(idx * step, 20) & 4095
            acc += table[ph] >> sh
        if kind == "speech":
            env = (table[_lsr(idx * ((4 << 32) // sample_rate), 20) & 4095] + 16384) >> 7
            acc = (acc * env) >> 8
        h = _splitmix64(idx ^ _i64(((seed + 7919 * c) << 40) & _M64))
        noise = (h & (2 * noise_lsb - 1)) - noise_lsb
        cols.append(acc + noise)
    if channels == 2:
        cols[1] = torch.div(cols[0] * 13, 16, rounding_mode="floor") + torch.div(cols[1], 4, rounding_mode="floor")
    out = torch.stack(cols, dim=1).clamp_(-32768, 32767).to(torch.int16)
    return out.reshape(-1)


def synth_pcm16_long(n: int, channels: int, sample_rate: int, seed: int, kind: str, noise_lsb: int, device,
                     chunk: int = 1 << 22) -> torch.Tensor:
    out = torch.empty(n * channels, dtype=torch.int16, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        out[s * channels:(s + m) * channels] = synth_pcm16(m, channels, sample_rate, seed, kind, noise_lsb, device, s, n)
    return out
