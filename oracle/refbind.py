"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libzen_ref.so, i.e. the
UNMODIFIED reference sources (/root/reference/libzen/hps.cu, core.cu and the
headers they include) behind the flat entry points of oracle/ref/ref_capi.cu.

Backend 0 = the reference GPU path (thrust + cuFFT + NPP; needs a GPU),
backend 1 = the reference CPU dataflow with its IPP calls served by
oracle/ref/ippstub/ipp.h (runs anywhere).

Only tests/, __graft_entry__.smoke() and bench.py's reference arm import this.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZEN_REF_SO selects another build of the reference, e.g. oracle/_ref/libzen_ref_norace.so (oracle/Makefile: ref_norace)
REF_SO = os.environ.get("ZEN_REF_SO") or os.path.join(_HERE, "_ref", "libzen_ref.so")

GPU, CPU = 0, 1
CAUSAL, ANTICAUSAL, FREQUENCY = 0, 1, 2
OUT_H, OUT_P, OUT_R = 1, 2, 4

FIELDS = {
    "s_mag": 0, "harmonic_matrix": 1, "percussive_matrix": 2, "harmonic_mask": 3,
    "percussive_mask": 4, "residual_mask": 5, "harmonic_out": 6, "percussive_out": 7,
    "residual_out": 8, "reciprocal": 9, "input": 10, "window": 11, "sliding_stft": 12,
}

_lib = None


def available() -> bool:
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(REF_SO)
        vp, ci, cf, cu, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint, ctypes.c_double
        L.ref_device_count.restype = ci
        L.ref_hpr_create.restype = vp
        L.ref_hpr_create.argtypes = [ci, cf, ci, cf, cu, ci, ci]
        for name in ("ref_hpr_destroy", "ref_hpr_use_sse", "ref_hpr_use_soft", "ref_hpr_reset"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = None
        L.ref_hpr_geometry.argtypes = [vp, vp]
        L.ref_hpr_process_next_hop.argtypes = [vp, vp]
        L.ref_hpr_get.argtypes = [vp, ci, vp]
        L.ref_hpr_get.restype = ci
        L.ref_hpr_run.argtypes = [vp, vp, ci, vp, vp, vp]
        L.ref_fakert_latency.argtypes = [ci, cf, ci, cf, ci, ci, vp, ci, ci, vp, vp]
        L.ref_fakert_latency.restype = ci
        L.ref_offline_process.argtypes = [ci, cf, ci, ci, cf, cf, ci, ci, vp, ctypes.c_long, vp, vp, vp]
        L.ref_offline_process.restype = cd
        L.ref_median_filter.argtypes = [ci, ci, ci, ci, ci, ci, vp, vp]
        L.ref_median_filter.restype = ci
        L.ref_box_filter.argtypes = [ci, ci, ci, ci, ci, vp, vp]
        L.ref_box_filter.restype = ci
        L.ref_median_filter_time.argtypes = [ci, ci, ci, ci, ci, ci]
        L.ref_median_filter_time.restype = cd
        L.ref_fft.argtypes = [ci, ci, vp, ci]
        L.ref_window.argtypes = [ci, ci, vp]
        _lib = L
        if L.ref_device_count() > 0:
            # Run one thrust kernel before any NPP/cuFFT call, as the reference's own
            # callers do (HPR's constructor fills its device_vectors first, hps.h:231-244).
            # On this toolkit a process whose first thrust kernel comes AFTER NPP calls
            # fails with "parallel_for failed: invalid device ordinal" (observed on the B200 box).
            L.ref_debug_thrust()
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


class RefHPR:
    """zen::internal::hps::HPR<B> (libzen/hps.h:152-322) driven hop by hop."""

    def __init__(self, backend, fs, hop, beta, flags, causality, copy_bord):
        self._h = lib().ref_hpr_create(backend, fs, hop, beta, flags, causality, int(copy_bord))
        if not self._h:
            raise ValueError("ZgException")
        g = (ctypes.c_double * 10)()
        lib().ref_hpr_geometry(self._h, g)
        self.fs, self.hop, self.nwin, self.nfft = g[0], int(g[1]), int(g[2]), int(g[3])
        self.beta, self.l_harm, self.l_perc, self.lag = g[4], int(g[5]), int(g[6]), int(g[7])
        self.stft_width, self.cola = int(g[8]), float(np.float32(g[9]))

    def close(self):
        if self._h:
            lib().ref_hpr_destroy(self._h)
            self._h = None

    __del__ = close

    def use_sse_filter(self):
        lib().ref_hpr_use_sse(self._h)

    def use_soft_mask(self):
        lib().ref_hpr_use_soft(self._h)

    def reset_buffers(self):
        lib().ref_hpr_reset(self._h)

    def process_next_hop(self, hop_in):
        a = np.ascontiguousarray(hop_in, dtype=np.float32)
        assert a.size == self.hop
        lib().ref_hpr_process_next_hop(self._h, _p(a))

    def get(self, name):
        which = FIELDS[name]
        n = self.stft_width * self.nfft
        if which in (6, 7, 8, 10, 11):
            n = self.nwin
        if which == 12:
            buf = np.empty(2 * n, dtype=np.float32)
            lib().ref_hpr_get(self._h, which, _p(buf))
            return buf.view(np.complex64).reshape(self.stft_width, self.nfft)
        buf = np.empty(n, dtype=np.float32)
        lib().ref_hpr_get(self._h, which, _p(buf))
        return buf if which in (6, 7, 8, 10, 11) else buf.reshape(self.stft_width, self.nfft)

    def run(self, audio, n_hops=None, want=(True, True, True)):
        a = np.ascontiguousarray(audio, dtype=np.float32)
        if n_hops is None:
            n_hops = a.size // self.hop
        outs = [np.zeros(n_hops * self.hop, dtype=np.float32) if w else None for w in want]
        lib().ref_hpr_run(self._h, _p(a), n_hops, _p(outs[0]), _p(outs[1]), _p(outs[2]))
        return outs


def fakert_latency(backend, fs, hop, beta, audio, n_hops, nocopybord=False, sse=False, soft=False, warm=True):
    a = np.ascontiguousarray(audio, dtype=np.float32)
    assert a.size >= n_hops * hop
    perc = np.zeros(n_hops * hop, dtype=np.float32)
    us = np.zeros(n_hops, dtype=np.float64)
    rc = lib().ref_fakert_latency(backend, fs, hop, beta, int(nocopybord), int(sse) | (int(soft) << 1),
                                  _p(a), n_hops, int(warm), _p(perc), _p(us))
    if rc != 0:
        raise ValueError("ZgException")
    return perc, us


def offline_process(backend, fs, hop_h, hop_p, beta_h, beta_p, audio, nocopybord=False, sse=False, soft=False):
    a = np.ascontiguousarray(audio, dtype=np.float32)
    outs = [np.zeros(a.size, dtype=np.float32) for _ in range(3)]
    ms = lib().ref_offline_process(backend, fs, hop_h, hop_p, beta_h, beta_p, int(nocopybord),
                                   int(sse) | (int(soft) << 1), _p(a), a.size, *[_p(o) for o in outs])
    if ms < 0:
        raise ValueError("ZgException")
    return outs, ms


def median_filter(backend, src, filter_len, direction, copy_bord, dst_init=None):
    s = np.ascontiguousarray(src, dtype=np.float32)
    T, F = s.shape
    d = np.zeros_like(s) if dst_init is None else np.ascontiguousarray(dst_init, dtype=np.float32).copy()
    rc = lib().ref_median_filter(backend, T, F, filter_len, direction, int(copy_bord), _p(s), _p(d))
    if rc != 0:
        raise ValueError("ZgException")
    return d


def box_filter(backend, src, filter_len, direction, dst_init=None):
    s = np.ascontiguousarray(src, dtype=np.float32)
    T, F = s.shape
    d = np.zeros_like(s) if dst_init is None else np.ascontiguousarray(dst_init, dtype=np.float32).copy()
    rc = lib().ref_box_filter(backend, T, F, filter_len, direction, _p(s), _p(d))
    if rc != 0:
        raise ValueError("ZgException")
    return d


def fft(backend, x, inverse=False):
    a = np.ascontiguousarray(x, dtype=np.complex64).copy()
    lib().ref_fft(backend, a.size, _p(a), int(inverse))
    return a


def window(n, sqrt=True):
    w = np.zeros(n, dtype=np.float32)
    lib().ref_window(0 if sqrt else 1, n, _p(w))
    return w
