"""TEST INFRASTRUCTURE - golden beat-tracker runs from the UNMODIFIED reference (demos/beat-tracking/BTrack.cpp +
OnsetDetection.cpp compiled into oracle/_ref/libbtrack_ref.so by `make -C oracle ref_btrack`, IPP FFTs served by
oracle/ref/ippstub).

    python oracle/ref/make_btrack_golden.py   -> tests/golden/btrack.npz

Per fixture (tests/btrack_inputs.py, regenerated from seeds by the tests): the onset detection samples the tracker
consumed, and per hop the cumulative score, the beat flag and the tempo estimate.  The two lookup tables of
BTrackPrecomputed.h are stored as well."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.btrack_inputs import CASES, make_input  # noqa: E402

L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libbtrack_ref.so"))
vp = ctypes.c_void_p
L.ref_btrack_run.argtypes = [ctypes.c_int, vp, ctypes.c_long, vp, vp, vp, vp]
L.ref_btrack_tables.argtypes = [vp, vp]
out = {}
r = np.zeros(128, np.float32)
t = np.zeros(41 * 41, np.float32)
L.ref_btrack_tables(r.ctypes.data, t.ctypes.data)
out["rayleigh"], out["transition"] = r, t.reshape(41, 41)
for name, fs, n_hops, kind, arg in CASES:
    x = make_input(fs, n_hops, kind, arg)
    odf = np.zeros(n_hops, np.float32)
    score = np.zeros(n_hops, np.float32)
    beat = np.zeros(n_hops, np.uint8)
    tempo = np.zeros(n_hops, np.float32)
    L.ref_btrack_run(int(fs), x.ctypes.data, n_hops, odf.ctypes.data, score.ctypes.data, beat.ctypes.data, tempo.ctypes.data)
    out[name + "_odf"], out[name + "_score"], out[name + "_beat"], out[name + "_tempo"] = odf, score, beat, tempo
    print(name, "beats", int(beat.sum()), "final tempo", float(tempo[-1]), "nan odf", int(np.isnan(odf).sum()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "btrack.npz"), **out)
