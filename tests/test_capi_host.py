"""The C-ABI library loads, exports every symbol include/zen_b200.h declares, and
its host-side logic (derived sizes, window, argument / geometry errors) matches
the oracle — no kernel is launched, so this runs on a box without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from zen_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zen_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zen_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(n for n in _lib.EXPORTS) == names


def test_version_and_device_count():
    L = _lib.lib()
    assert b"sm_100a" in L.zen_b200_version()
    assert L.zen_device_count() >= 0


@pytest.mark.parametrize("fs,hop,causal", [(44100.0, 256, 1), (44100.0, 512, 0), (44100.0, 1024, 1), (44100.0, 2048, 0),
                                           (44100.0, 4096, 1), (48000.0, 256, 0), (48000.0, 1024, 1), (22050.0, 512, 1)])
def test_geometry_matches_oracle(oracle, fs, hop, causal):
    g = _lib.ZenGeometry()
    assert _lib.lib().zen_hpr_geometry(fs, hop, causal, ctypes.byref(g)) == _lib.ZEN_OK
    o = oracle.geometry(fs, hop, bool(causal))
    assert (g.hop, g.nwin, g.nfft, g.l_harm, g.l_perc, g.lag, g.stft_width) == (o.hop, o.nwin, o.nfft, o.l_harm, o.l_perc, o.lag, o.stft_width)
    assert np.float32(g.cola_factor) == np.float32(o.cola)


@pytest.mark.parametrize("n", [512, 1024, 2048, 8192])
def test_window_bit_identical_to_oracle(oracle, n):
    for sqrt in (True, False):
        w = np.zeros(n, dtype=np.float32)
        assert _lib.lib().zen_window(0 if sqrt else 1, n, w.ctypes.data) == _lib.ZEN_OK
        assert np.array_equal(w, oracle.window(n, sqrt=sqrt))


def test_argument_and_geometry_errors_without_compute():
    L = _lib.lib()
    # ZgException cases are decided on the host before anything is launched
    assert L.zen_median_filter(9, 9, 10, _lib.TIME_CAUSAL, 0, 1, 1, None) == _lib.ZEN_ERR_GEOMETRY
    assert L.zen_median_filter(9, 9, 10, _lib.FREQUENCY, 1, 1, 1, None) == _lib.ZEN_ERR_GEOMETRY
    assert L.zen_box_filter(9, 9, 10, _lib.TIME_ANTICAUSAL, 1, 1, None) == _lib.ZEN_ERR_GEOMETRY
    assert L.zen_median_filter(9, 9, 3, 0, 0, None, None, None) == _lib.ZEN_ERR_ARG
    assert L.zen_fft_c2c(1000, 1, 0, None) == _lib.ZEN_ERR_UNSUPPORTED
    assert L.zen_fft_c2c(1 << 17, 1, 0, None) == _lib.ZEN_ERR_UNSUPPORTED
    a = np.zeros(8, np.float32)
    assert L.zen_offline_process(44100.0, 4096, 300, 2.0, 2.0, 0, a.ctypes.data, 8, a.ctypes.data, a.ctypes.data,
                                 a.ctypes.data) == _lib.ZEN_ERR_GEOMETRY     # hps.cu:33-36


def test_no_cpu_fallback():
    """without a CUDA device every computing entry point must fail loudly"""
    L = _lib.lib()
    if L.zen_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert L.zen_hpr_create(ctypes.byref(h), 44100.0, 1024, 2.5, 7, 0, 1) == _lib.ZEN_ERR_CUDA
    b = ctypes.c_void_p()
    assert L.zen_hpr_batch_create(ctypes.byref(b), 44100.0, 1024, 2.5, 2, 0, 0, 1, 1) == _lib.ZEN_ERR_CUDA
    a = np.zeros(4096 * 4, np.float32)
    assert L.zen_offline_process(44100.0, 4096, 256, 2.0, 2.0, 0, a.ctypes.data, a.size, a.ctypes.data, a.ctypes.data,
                                 a.ctypes.data) == _lib.ZEN_ERR_CUDA
    from zen_b200 import hps
    with pytest.raises(_lib.ZenCudaError):
        hps.HPRRealtime(44100.0, 1024, 2.5, hps.OUTPUT_PERCUSSIVE)


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/"""
    banned = re.compile(r"(import\s+oracle|from\s+oracle|oraclebind|refbind|libzen_oracle|libzen_ref|hpr_oracle\.h|#include\s+[<\"][^>\"]*oracle)")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zen_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not banned.search(txt), os.path.join(dirpath, f)
