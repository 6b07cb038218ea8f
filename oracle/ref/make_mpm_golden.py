"""TEST INFRASTRUCTURE - golden pitches from the UNMODIFIED reference consumer (demos/pitch-tracking/pitch.cpp compiled
into oracle/_ref/libmpm_ref.so by `make -C oracle ref_mpm`, IPP FFT served by oracle/ref/ippstub).

    python oracle/ref/make_mpm_golden.py   -> tests/golden/mpm_pitch.npz

Inputs are regenerated from seeds by the tests (tones with harmonics, the synthetic music-like signal, noise, silence)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.mpm_inputs import CASES, make_input  # noqa: E402

L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmpm_ref.so"))
L.ref_mpm_pitch.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_float]
L.ref_mpm_pitch.restype = ctypes.c_float
out = {}
for name, n, fs, kind, arg in CASES:
    x = make_input(n, fs, kind, arg)
    out[name] = np.float32(L.ref_mpm_pitch(x.ctypes.data, n, fs))
    print(name, out[name])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mpm_pitch.npz"), **out)
