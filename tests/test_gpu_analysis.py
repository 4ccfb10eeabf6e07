"""Waveform peaks of libflo::encode()'s analysis metadata (libflo/src/core/analysis.rs:38-119):
the CUDA path through the C ABI against the oracle's restatement, bit for bit (parity unpinned: see
oracle/flo_analysis.py)."""
import numpy as np
import pytest

from helpers import oracle, pcm16_to_f32, synth_pcm16  # puts oracle/ on the path
import flo_analysis  # noqa: E402  (oracle/: the checker)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fa():
    import flo_b200
    return flo_b200.analysis


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("ch,sr,pps,n", [
    (1, 44100, 50, 44100 * 3 + 17), (2, 44100, 50, 2 * (44100 * 3) + 1), (2, 48000, 60, 2 * 48000 * 2),
    (1, 22050, 7, 22050 * 2 + 5), (2, 44100, 10, 6), (1, 8000, 50, 8000 * 5), (3, 44100, 50, 3 * 20000 + 2),
    (6, 48000, 25, 6 * 30000 + 4), (1, 5, 2, 10), (1, 100, 1000, 250), (2, 100, 1000, 251),
])
def test_peaks_match_the_oracle_bit_for_bit(fa, ch, sr, pps, n):
    rng = np.random.default_rng(0xA11 + ch + sr + pps)
    frames = -(-n // ch)
    x = pcm16_to_f32(synth_pcm16(frames, ch, max(sr, 8000), seed=0xA12 + ch, kind="multitone", noise_lsb=64))[:n].copy()
    x *= np.float32(0.25) + rng.random(n, dtype=np.float32) * np.float32(0.01)      # full f32 mantissas
    if ch > 2:
        x -= np.float32(0.1)                                                        # the many-channel arm takes signed means
    want = flo_analysis.extract_waveform_peaks(x, ch, sr, pps)
    got = fa.extract_waveform_peaks(x, ch, sr, pps)
    assert got.channels == ch and got.peaks_per_second == pps
    assert got.peaks.size == fa.peaks_count(n, sr, ch, pps) == want.size
    assert same_bits(got.peaks, want)
    if want.size:
        assert float(np.nanmax(got.peaks)) <= 1.0 and float(np.nanmin(got.peaks)) >= 0.0        # analysis_tests.rs:13-15


def test_reference_test_inputs(fa):
    samples = np.array([0.5, -0.3, 0.8, -0.2, 0.1, -0.9], np.float32)              # analysis_tests.rs:5, :21, :60
    for ch in (1, 2):
        w1 = fa.extract_waveform_peaks(samples, ch, 44100, 10)
        w2 = fa.extract_waveform_peaks(samples, ch, 44100, 10)
        assert w1.peaks.size > 0 and same_bits(w1.peaks, w2.peaks) and w1.channels == ch and w1.peaks_per_second == 10
        assert same_bits(w1.peaks, flo_analysis.extract_waveform_peaks(samples, ch, 44100, 10))
    assert fa.extract_waveform_peaks(np.zeros(0, np.float32), 1, 44100, 10).peaks.size == 0   # analysis_tests.rs:34-41


def test_corners(fa):
    import flo_b200
    assert fa.extract_waveform_peaks([1.0, 2.0], 1, 44100, 0).peaks.size == 0       # peaks_per_second = 0: no window
    assert fa.extract_waveform_peaks([0.0, 0.0, 0.0], 1, 1, 1).peaks.tolist() == [0.0, 0.0, 0.0]
    nan = float("nan")
    assert fa.extract_waveform_peaks([nan, 0.5, nan, nan], 1, 2, 1).peaks.tolist() == [1.0, 0.0]
    x = np.array([0.5, -0.25, 0.25, 0.75, 1.0, 0.0, -0.5], np.float32)
    assert same_bits(fa.extract_waveform_peaks(x, 2, 2, 1).peaks, flo_analysis.extract_waveform_peaks(x, 2, 2, 1))
    for ch, sr in ((0, 44100), (1, 0)):
        with pytest.raises(flo_b200.FloError):
            fa.extract_waveform_peaks([1.0], ch, sr, 10)


def test_device_entry_on_ten_minutes(fa):
    import torch
    sr, ch, pps = 44100, 2, 50
    n = sr * ch * 600
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(7)
    x = (torch.rand(n, device=dev, generator=g) - 0.5) * torch.linspace(0.1, 1.0, n, device=dev)
    cap = fa.peaks_count(n, sr, ch, pps)
    assert cap == 600 * pps
    out = torch.empty(cap, dtype=torch.float32, device=dev)
    got_n = fa.extract_waveform_peaks_device(x.data_ptr(), n, ch, sr, pps, out.data_ptr(), cap)
    assert got_n == cap
    host = x.cpu().numpy()
    want = flo_analysis.extract_waveform_peaks(host, ch, sr, pps)
    assert same_bits(out.cpu().numpy(), want)
    with pytest.raises(Exception):
        fa.extract_waveform_peaks_device(x.data_ptr(), n, ch, sr, pps, out.data_ptr(), cap - 1)


# ---- EBU R128 integrated loudness (libflo/src/core/ebu_r128.rs:182-313) ------------------------------------------
LUFS_TOL = 1e-9      # LU; the device chains 100 ms hops instead of one serial recurrence per channel


@pytest.mark.parametrize("ch,sr,secs,kind", [
    (1, 44100, 3.0, "multitone"), (2, 44100, 5.3, "speech"), (2, 48000, 2.0, "sweep"), (1, 8000, 7.0, "speech"),
    (6, 48000, 1.5, "multitone"), (2, 96000, 1.2, "sweep"), (1, 22050, 0.35, "multitone"), (2, 44100, 0.05, "multitone"),
    (3, 12345, 2.2, "multitone"),
])
def test_integrated_loudness_matches_the_oracle(fa, ch, sr, secs, kind):
    n = int(sr * secs)
    x = pcm16_to_f32(synth_pcm16(n, ch, max(sr, 8000), seed=0xB00 + ch + sr, kind=kind, noise_lsb=32))
    x = (x * np.float32(0.6)).astype(np.float32)
    want = oracle.r128_integrated_lufs(x, ch, sr)
    got = fa.compute_ebu_r128_loudness(x, ch, sr).integrated_lufs
    assert abs(got - want) <= LUFS_TOL, (got, want)
    assert np.float32(got) == np.float32(want)                                          # what encode() stores (lib.rs:262-265)
    assert abs(fa.compute_ebu_r128_loudness(x[:-1], ch, sr).integrated_lufs - oracle.r128_integrated_lufs(x[:-1], ch, sr)) <= LUFS_TOL


def test_loudness_corners_and_calibration(fa):
    assert fa.compute_ebu_r128_loudness(np.zeros(0, np.float32), 1, 44100).integrated_lufs == -23.0      # loudness_tests.rs:4-12
    assert fa.compute_ebu_r128_loudness(np.zeros(44100, np.float32), 1, 44100).integrated_lufs == -23.0  # :15-23
    assert fa.compute_ebu_r128_loudness(np.ones(10, np.float32), 0, 44100).integrated_lufs == -23.0      # ebu_r128.rs:187
    assert fa.compute_ebu_r128_loudness(np.ones(1, np.float32), 2, 44100).integrated_lufs == -23.0       # no whole frame
    import flo_b200
    x1k = pcm16_to_f32(synth_pcm16(3000, 1, 8000, seed=3))
    assert oracle.r128_integrated_lufs(x1k, 1, 1000) == float("inf")                   # unstable shelf below ~3.4 kHz: overflow
    with pytest.raises(flo_b200.FloError):
        fa.compute_ebu_r128_loudness(x1k, 1, 1000)
    sr = 48000
    s997 = np.sin(2 * np.pi * 997 * np.arange(sr * 5) / sr).astype(np.float32)
    assert abs(fa.compute_ebu_r128_loudness(s997, 1, sr).integrated_lufs - (-3.01)) < 0.01               # BS.1770 calibration
    loud = (0.5 * np.sin(2 * np.pi * 440 * np.arange(44100 * 4) / 44100)).astype(np.float32)
    both = np.concatenate([loud, loud * np.float32(10 ** -1.5), np.zeros(44100, np.float32)])           # both gates at work
    assert abs(fa.compute_ebu_r128_loudness(both, 1, 44100).integrated_lufs - oracle.r128_integrated_lufs(both, 1, 44100)) <= LUFS_TOL


def test_loudness_device_entry_on_ten_minutes(fa):
    import torch
    sr, ch = 44100, 2
    n = sr * ch * 600
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11)
    x = (torch.rand(n, device=dev, generator=g) - 0.5) * (0.05 + 0.5 * torch.sin(torch.linspace(0, 40, n, device=dev)) ** 2)
    got = fa.integrated_loudness_device(x.data_ptr(), n, ch, sr)
    want = oracle.r128_integrated_lufs(x.cpu().numpy(), ch, sr)
    assert abs(got - want) <= LUFS_TOL, (got, want)
