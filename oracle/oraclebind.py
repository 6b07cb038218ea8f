"""TEST INFRASTRUCTURE — ctypes binding of oracle/libzen_oracle.so (the plain-C
restatement in oracle/hpr_oracle.c).  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libzen_oracle.so")

GEOM_GPU, GEOM_CPU = 0, 1
CAUSAL, ANTICAUSAL, FREQUENCY = 0, 1, 2
OUT_H, OUT_P, OUT_R = 1, 2, 4
FIELDS = {
    "s_mag": 0, "harmonic_matrix": 1, "percussive_matrix": 2, "harmonic_mask": 3,
    "percussive_mask": 4, "residual_mask": 5, "harmonic_out": 6, "percussive_out": 7,
    "residual_out": 8, "reciprocal": 9, "input": 10, "window": 11, "sliding_stft": 12,
}


class ZoGeom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("hop", "nwin", "nfft", "l_harm", "l_perc", "lag", "stft_width")] + [
        ("cola", ctypes.c_float)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hpr_oracle.c")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            build()
        L = ctypes.CDLL(ORACLE_SO)
        vp, ci, cf, cu, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint, ctypes.c_long
        L.zo_geometry.argtypes = [cf, ci, ci, ctypes.POINTER(ZoGeom)]
        L.zo_window.argtypes = [ci, ci, vp]
        L.zo_median_filter.argtypes = [ci, ci, ci, ci, ci, ci, vp, vp]
        L.zo_box_filter.argtypes = [ci, ci, ci, ci, ci, vp, vp]
        L.zo_fft.argtypes = [ci, vp, ci]
        L.zo_hpr_create.restype = vp
        L.zo_hpr_create.argtypes = [ci, cf, ci, cf, cu, ci, ci]
        for n in ("zo_hpr_destroy", "zo_hpr_use_sse_filter", "zo_hpr_use_soft_mask", "zo_hpr_reset_buffers"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = None
        L.zo_hpr_geom.argtypes = [vp, ctypes.POINTER(ZoGeom)]
        L.zo_hpr_process_next_hop.argtypes = [vp, vp]
        L.zo_hpr_get.argtypes = [vp, ci, vp]
        L.zo_hpr_run.argtypes = [vp, vp, ci, vp, vp, vp]
        L.zo_offline_process.argtypes = [ci, cf, ci, ci, cf, cf, ci, ci, vp, cl, vp, vp, vp]
        L.zo_mpm_pitch.argtypes = [vp, ci, cf, vp]
        L.zo_mpm_pitch.restype = cf
        L.zo_onset_csd.argtypes = [vp, cl, vp]
        L.zo_onset_csd.restype = None
        L.zo_onset_window.argtypes = [vp]
        L.zo_onset_window.restype = None
        L.zo_fakert_n_chunks.argtypes = [cl, cl]
        L.zo_fakert_n_chunks.restype = cl
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


def geometry(fs, hop, causal):
    g = ZoGeom()
    lib().zo_geometry(fs, hop, int(causal), ctypes.byref(g))
    return g


def window(n, sqrt=True):
    w = np.zeros(n, dtype=np.float32)
    lib().zo_window(0 if sqrt else 1, n, _p(w))
    return w


def median_filter(geom, src, filter_len, direction, copy_bord, dst_init=None):
    s = np.ascontiguousarray(src, dtype=np.float32)
    T, F = s.shape
    d = np.zeros_like(s) if dst_init is None else np.ascontiguousarray(dst_init, dtype=np.float32).copy()
    if lib().zo_median_filter(geom, T, F, filter_len, direction, int(copy_bord), _p(s), _p(d)) != 0:
        raise ValueError("ZgException")
    return d


def box_filter(geom, src, filter_len, direction, dst_init=None):
    s = np.ascontiguousarray(src, dtype=np.float32)
    T, F = s.shape
    d = np.zeros_like(s) if dst_init is None else np.ascontiguousarray(dst_init, dtype=np.float32).copy()
    if lib().zo_box_filter(geom, T, F, filter_len, direction, _p(s), _p(d)) != 0:
        raise ValueError("ZgException")
    return d


def fft(x, inverse=False):
    a = np.ascontiguousarray(x, dtype=np.complex64).copy()
    lib().zo_fft(a.size, _p(a), int(inverse))
    return a


class OracleHPR:
    """Restatement of zen::internal::hps::HPR<B> (libzen/hps.h:152-322)."""

    def __init__(self, geom, fs, hop, beta, flags, causality, copy_bord):
        self._h = lib().zo_hpr_create(geom, fs, hop, beta, flags, causality, int(copy_bord))
        if not self._h:
            raise ValueError("ZgException")
        g = ZoGeom()
        lib().zo_hpr_geom(self._h, ctypes.byref(g))
        self.beta = beta
        self.hop, self.nwin, self.nfft = g.hop, g.nwin, g.nfft
        self.l_harm, self.l_perc, self.lag, self.stft_width, self.cola = g.l_harm, g.l_perc, g.lag, g.stft_width, g.cola

    def close(self):
        if self._h:
            lib().zo_hpr_destroy(self._h)
            self._h = None

    __del__ = close

    def use_sse_filter(self):
        lib().zo_hpr_use_sse_filter(self._h)

    def use_soft_mask(self):
        lib().zo_hpr_use_soft_mask(self._h)

    def reset_buffers(self):
        lib().zo_hpr_reset_buffers(self._h)

    def process_next_hop(self, hop_in):
        a = np.ascontiguousarray(hop_in, dtype=np.float32)
        assert a.size == self.hop
        lib().zo_hpr_process_next_hop(self._h, _p(a))

    def get(self, name):
        which = FIELDS[name]
        n = self.stft_width * self.nfft
        if which in (6, 7, 8, 10, 11):
            buf = np.empty(self.nwin, dtype=np.float32)
            lib().zo_hpr_get(self._h, which, _p(buf))
            return buf
        if which == 12:
            buf = np.empty(2 * n, dtype=np.float32)
            lib().zo_hpr_get(self._h, which, _p(buf))
            return buf.view(np.complex64).reshape(self.stft_width, self.nfft)
        buf = np.empty(n, dtype=np.float32)
        lib().zo_hpr_get(self._h, which, _p(buf))
        return buf.reshape(self.stft_width, self.nfft)

    def run(self, audio, n_hops=None, want=(True, True, True)):
        a = np.ascontiguousarray(audio, dtype=np.float32)
        if n_hops is None:
            n_hops = a.size // self.hop
        outs = [np.zeros(n_hops * self.hop, dtype=np.float32) if w else None for w in want]
        lib().zo_hpr_run(self._h, _p(a), n_hops, _p(outs[0]), _p(outs[1]), _p(outs[2]))
        return outs


def offline_process(geom, fs, hop_h, hop_p, beta_h, beta_p, audio, nocopybord=False, sse=False, soft=False):
    a = np.ascontiguousarray(audio, dtype=np.float32)
    outs = [np.zeros(a.size, dtype=np.float32) for _ in range(3)]
    rc = lib().zo_offline_process(geom, fs, hop_h, hop_p, beta_h, beta_p, int(nocopybord),
                                  int(sse) | (int(soft) << 1), _p(a), a.size, *[_p(o) for o in outs])
    if rc != 0:
        raise ValueError("ZgException")
    return outs


def fakert_n_chunks(size, hop):
    return int(lib().zo_fakert_n_chunks(size, hop))


def mpm_pitch(audio, sample_rate, want_nsdf=False):
    """demos/pitch-tracking/pitch.cpp MPM::pitch on one buffer: pitch in Hz or -1 (and the autocorrelation it picked from)"""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    nsdf = np.zeros(a.size, dtype=np.float32) if want_nsdf else None
    p = float(lib().zo_mpm_pitch(_p(a), a.size, sample_rate, _p(nsdf)))
    return (p, nsdf) if want_nsdf else p


def onset_csd(audio):
    """demos/beat-tracking/OnsetDetection.cpp calculate_sample over consecutive 256-sample hops: one ODF sample per hop"""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    n_hops = a.size // 256
    out = np.zeros(n_hops, dtype=np.float32)
    lib().zo_onset_csd(_p(a), n_hops, _p(out))
    return out


def onset_window():
    w = np.zeros(512, dtype=np.float32)
    lib().zo_onset_window(_p(w))
    return w
