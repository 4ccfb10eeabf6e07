"""C-ABI checks that need no GPU: the library loads, exports every symbol include/flo_b200.h declares,
fails loudly without a device, and its host-only layout helper agrees with the oracle's output sizes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, oracle, pcm16_to_f32, synth_pcm16


def header_functions():
    src = open(os.path.join(ROOT, "include", "flo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from flo_b200 import _lib
    L = _lib.lib()
    declared = header_functions()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/flo_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared


def test_no_cpu_fallback_without_device():
    import torch
    from flo_b200 import _lib
    import flo_b200
    L = _lib.lib()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert L.flo_device_count() == 0
    h = C.c_void_p()
    assert L.flo_ctx_create(0, C.byref(h)) != 0 and not h.value
    assert b"no CPU fallback" in L.flo_last_error()
    with pytest.raises(flo_b200.FloError):
        flo_b200.Encoder(44100, 2, 16).encode(np.zeros(16, np.float32), b"")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "flo_oracle" not in text and "flo_ref_" not in text, f"{f} references the oracle"


def test_output_bound_covers_oracle_sizes():
    from flo_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    cases = [(5000, 1, 8000), (44100 + 7, 2, 44100), (3001, 3, 1000), (0, 2, 44100), (1, 1, 8000)]
    tracks = (_lib.Track * len(cases))()
    total = 0
    for i, (n, ch, sr) in enumerate(cases):
        pcm = rng.integers(-32768, 32768, n * ch, dtype=np.int64).astype(np.int16)     # white noise: worst case (raw)
        out = oracle.encode(pcm16_to_f32(pcm), sr, ch, 16, 5, b"meta")
        total += len(out)
        tracks[i].samples = 16 if n else None
        tracks[i].n_interleaved = n * ch
        tracks[i].sample_rate = sr
        tracks[i].channels = ch
        tracks[i].bit_depth = 16
        tracks[i].meta = 16
        tracks[i].meta_len = 4
    bound = L.flo_output_bound(tracks, len(cases))
    assert bound >= total
    assert bound <= total + 64 + 60 * sum(-(-n // sr) * ch for n, ch, sr in cases) + 16


def test_argument_errors_mirror_reference_panics():
    from flo_b200 import _lib
    L = _lib.lib()
    t = (_lib.Track * 1)()
    t[0].samples = 16
    t[0].n_interleaved = 8
    t[0].sample_rate = 44100
    t[0].channels = 0                   # encoder.rs:48: division by zero panic in the reference
    assert L.flo_output_bound(t, 1) == 0 and b"channels == 0" in L.flo_last_error()
    t[0].channels = 2
    t[0].sample_rate = 0
    assert L.flo_output_bound(t, 1) == 0 and b"sample_rate == 0" in L.flo_last_error()


def test_python_mirror_of_encoder_interface():
    import flo_b200
    e = flo_b200.Encoder()
    assert (e.sample_rate, e.channels, e.bit_depth, e.compression_level) == (44100, 1, 16, 5)   # Default + level 5
    assert flo_b200.Encoder(48000, 2, 24).with_compression(200).compression_level == 9          # level.min(9)
    with pytest.raises(flo_b200.FloError):
        flo_b200.Encoder(44100, 256, 16)


def test_rust_shim_binds_exactly_the_header_symbols():
    """rust/flo-b200 cannot be compiled here (no cargo); keep its extern block mechanically in sync."""
    src = open(os.path.join(ROOT, "rust", "flo-b200", "src", "ffi.rs")).read()
    bound = sorted(set(re.findall(r"pub fn (flo_[a-z0-9_]+)\s*\(", src)))
    assert bound == header_functions()
