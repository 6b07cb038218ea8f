"""TEST INFRASTRUCTURE - golden onset detection function samples from the UNMODIFIED reference consumer
(demos/beat-tracking/OnsetDetection.cpp compiled into oracle/_ref/libodf_ref.so by `make -C oracle ref_odf`, IPP FFT
served by oracle/ref/ippstub, compile-time window by the vendored gcem).

    python oracle/ref/make_odf_golden.py   -> tests/golden/odf_csd.npz

Inputs are regenerated from seeds by the tests (tests/odf_inputs.py)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.odf_inputs import CASES, make_input  # noqa: E402

L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libodf_ref.so"))
L.ref_odf_run.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
L.ref_odf_window.argtypes = [ctypes.c_void_p]
out = {}
w = np.zeros(512, np.float32)
L.ref_odf_window(w.ctypes.data)
out["window"] = w
for name, n_hops, kind, arg in CASES:
    x = make_input(n_hops, kind, arg)
    y = np.zeros(n_hops, np.float32)
    L.ref_odf_run(x.ctypes.data, n_hops, y.ctypes.data)
    out[name] = y
    print(name, y[:4], float(y.max()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "odf_csd.npz"), **out)
