"""TEST INFRASTRUCTURE - golden vectors of the command line's on-disk sample format, produced by the UNMODIFIED
libnyquist conversion code (oracle/_ref/libnyq_ref.so, built by `make -C oracle ref_nyq` from
/root/reference/vendor/libnyquist/src/Common.cpp) and the peak normalisation of zen/offline.h:180-192.

    python oracle/ref/make_pcm_golden.py        -> tests/golden/nyq_pcm.npz

Pins oracle/np_model.py (tests/test_pcm.py, CPU) and csrc/pcm.cu (GPU)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libnyq_ref.so"))
vp, cl = ctypes.c_void_p, ctypes.c_long
L.nyq_pcm16_to_float.argtypes = [vp, vp, cl]
L.nyq_stereo_to_mono.argtypes = [vp, vp, cl]
L.nyq_float_to_pcm16.argtypes = [vp, vp, cl]
L.zen_cli_normalize_encode.argtypes = [vp, vp, cl]
L.zen_cli_normalize_encode.restype = ctypes.c_float


def decode(pcm):
    out = np.empty(pcm.size, np.float32)
    L.nyq_pcm16_to_float(pcm.ctypes.data, out.ctypes.data, pcm.size)
    return out


def fold(inter):
    out = np.empty(inter.size // 2, np.float32)
    L.nyq_stereo_to_mono(inter.ctypes.data, out.ctypes.data, inter.size)
    return out


def encode(x):
    out = np.empty(x.size, np.int16)
    L.nyq_float_to_pcm16(x.ctypes.data, out.ctypes.data, x.size)
    return out


def norm_encode(x):
    out = np.empty(x.size, np.int16)
    pk = L.zen_cli_normalize_encode(x.ctypes.data, out.ctypes.data, x.size)
    return out, np.float32(pk)


rng = np.random.default_rng(2024)
g = {}
all16 = np.arange(-32768, 32768, dtype=np.int16)
g["decode_all_int16"] = decode(all16)                                  # every PCM16 code
st = rng.integers(-32768, 32768, 6000).astype(np.int16)
st[:6] = [100, 300, -32768, -32768, 32767, -32767]
g["stereo_pcm"] = st
g["stereo_mono"] = fold(decode(st))
# encode without normalisation: halfway cases, +-1, values just inside / outside a rounding boundary
e = np.concatenate([np.array([0.0, 1.0, -1.0, 0.5 / 32767, -0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -2.5 / 32767, 0.49999997 / 32767,
                              0.99998474, -0.99998474, 1e-9, -1e-9], np.float32),
                    rng.uniform(-1, 1, 6000).astype(np.float32)])
g["encode_in"] = e
g["encode_out"] = encode(e)
# the command line's normalise + encode on signals of different scale; the peak may be a negative sample
for i, scale in enumerate((1e-3, 0.7, 1.0, 37.5, 29998.0)):
    x = (rng.standard_normal(6000) * scale).astype(np.float32)
    if i == 1:
        x[1234] = -np.abs(x).max() * 2
    q, pk = norm_encode(x)
    g["norm%d_in" % i], g["norm%d_out" % i], g["norm%d_peak" % i] = x, q, pk
# silence: the reference divides 0 by 0 and converts NaN; record what this build of it (x86-64, glibc) writes
q, pk = norm_encode(np.zeros(64, np.float32))
g["silence_out"], g["silence_peak"] = q, pk
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "nyq_pcm.npz")
np.savez_compressed(out, **g)
print(out, "%.0f KB" % (os.path.getsize(out) / 1024), "silence ->", set(q.tolist()), "peak", pk)
