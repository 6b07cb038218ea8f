"""Host-side mirror of the libzen API for the HPR path, over the C ABI.

Same names, argument meaning and error behaviour as the reference classes
(libzen/libzen/hps.h, libzen/libzen/io.h, libzen/mfilt.h, box.h, fftw.h, win.h)
so the parity tests read like the reference's own tests.  Device memory is
held in torch CUDA tensors; every computation happens in libzen_b200.so.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (OUTPUT_HARMONIC, OUTPUT_PERCUSSIVE, OUTPUT_RESIDUAL, ZgException,  # noqa: F401
                   check)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.ZenCudaError("zen_b200 needs a CUDA device; there is no CPU fallback")
    return torch


def _dptr(x):
    """device pointer of a torch tensor, or pass an int address through"""
    if isinstance(x, int):
        return x
    return x.data_ptr()


class MedianFilterDirection:
    TimeCausal, TimeAnticausal, Frequency = 0, 1, 2


def hpr_geometry(fs, hop, causal):
    g = _lib.ZenGeometry()
    check(_lib.lib().zen_hpr_geometry(fs, hop, int(causal), ctypes.byref(g)), "zen_hpr_geometry")
    return g


def window(n, sqrt=True):
    """Window<T> (libzen/win.h:21-53)."""
    w = np.zeros(n, dtype=np.float32)
    check(_lib.lib().zen_window(0 if sqrt else 1, n, w.ctypes.data), "zen_window")
    return w


class IOGPU:
    """zen::io::IOGPU (libzen/libzen/io.h:16-81): mapped pinned in/out buffers."""

    def __init__(self, size):
        self._io = _lib.ZenIO()
        check(_lib.lib().zen_io_alloc(ctypes.byref(self._io), size), "zen_io_alloc")
        self.size = size
        self.host_in = np.ctypeslib.as_array(ctypes.cast(self._io.host_in, ctypes.POINTER(ctypes.c_float)), (size,))
        self.host_out = np.ctypeslib.as_array(ctypes.cast(self._io.host_out, ctypes.POINTER(ctypes.c_float)), (size,))
        self.device_in = self._io.device_in
        self.device_out = self._io.device_out

    def close(self):
        if self._io is not None:
            self.host_in = self.host_out = None
            _lib.lib().zen_io_free(ctypes.byref(self._io))
            self._io = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class PinnedArray:
    """[rows, cols] array (float32, or int16 for PCM16 payloads) in page-locked host memory (zen_host_alloc)."""

    def __init__(self, rows, cols, dtype=np.float32):
        dt = np.dtype(dtype)
        ct = {np.dtype(np.float32): ctypes.c_float, np.dtype(np.int16): ctypes.c_int16}[dt]
        self.nbytes = rows * cols * dt.itemsize
        self.ptr = _lib.lib().zen_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("zen_host_alloc(%d bytes) failed" % self.nbytes)
        self.array = np.ctypeslib.as_array(ctypes.cast(self.ptr, ctypes.POINTER(ct)), (rows, cols))

    def close(self):
        if self.ptr:
            self.array = None
            _lib.lib().zen_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class MedianFilterGPU:
    """libzen/mfilt.h:33-268.  filter(src, dst) on time x freq float32 CUDA tensors."""

    def __init__(self, time, frequency, filter_len, direction, copy_bord=False):
        if ((direction in (0, 1) and filter_len > time) or (direction == 2 and filter_len > frequency)):
            raise ZgException("median filter bigger than matrix dimension")
        self.time, self.frequency, self.filter_len = time, frequency, filter_len
        self.mydir, self.copy_bord = direction, bool(copy_bord)

    def filter(self, src, dst):
        check(_lib.lib().zen_median_filter(self.time, self.frequency, self.filter_len, self.mydir, int(self.copy_bord),
                                           _dptr(src), _dptr(dst), None), "zen_median_filter")


class BoxFilterGPU:
    """libzen/box.h:30-215."""

    def __init__(self, time, frequency, filter_len, direction):
        if ((direction in (0, 1) and filter_len > time) or (direction == 2 and filter_len > frequency)):
            raise ZgException("box filter bigger than matrix dimension")
        self.time, self.frequency, self.filter_len, self.mydir = time, frequency, filter_len, direction

    def filter(self, src, dst):
        check(_lib.lib().zen_box_filter(self.time, self.frequency, self.filter_len, self.mydir, _dptr(src), _dptr(dst), None),
              "zen_box_filter")


class FFTC2CWrapperGPU:
    """libzen/fftw.h:20-49: owns fft_vec (complex64 CUDA tensor), in-place forward()/backward()."""

    def __init__(self, nfft):
        torch = _torch()
        self.nfft = nfft
        self.fft_vec = torch.zeros(nfft, dtype=torch.complex64, device="cuda")

    def forward(self):
        check(_lib.lib().zen_fft_c2c(self.nfft, self.fft_vec.data_ptr(), 0, None), "zen_fft_c2c")

    def backward(self):
        check(_lib.lib().zen_fft_c2c(self.nfft, self.fft_vec.data_ptr(), 1, None), "zen_fft_c2c")


class HPR:
    """zen::internal::hps::HPR<Backend::GPU> (libzen/hps.h:152-322)."""

    def __init__(self, fs, hop, beta, output_flags, causality, copy_bord):
        h = ctypes.c_void_p()
        check(_lib.lib().zen_hpr_create(ctypes.byref(h), fs, hop, beta, output_flags, causality, int(copy_bord)), "zen_hpr_create")
        self._h = h
        g = _lib.ZenGeometry()
        check(_lib.lib().zen_hpr_get_geometry(self._h, ctypes.byref(g)), "zen_hpr_get_geometry")
        self.fs, self.hop, self.nwin, self.nfft, self.beta = fs, g.hop, g.nwin, g.nfft, beta
        self.l_harm, self.l_perc, self.lag, self.stft_width = g.l_harm, g.l_perc, g.lag, g.stft_width
        self.COLA_factor = g.cola_factor
        self.output_harmonic = bool(output_flags & OUTPUT_HARMONIC)
        self.output_percussive = bool(output_flags & OUTPUT_PERCUSSIVE)
        self.output_residual = bool(output_flags & OUTPUT_RESIDUAL)
        self.use_sse = self.soft_mask = False

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().zen_hpr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def use_sse_filter(self):
        self.use_sse = True
        check(_lib.lib().zen_hpr_use_sse_filter(self._h), "use_sse_filter")

    def use_soft_mask(self):
        self.soft_mask = True
        check(_lib.lib().zen_hpr_use_soft_mask(self._h), "use_soft_mask")

    def reset_buffers(self):
        check(_lib.lib().zen_hpr_reset_buffers(self._h), "reset_buffers")

    def process_next_hop(self, in_hop):
        check(_lib.lib().zen_hpr_process_next_hop(self._h, _dptr(in_hop)), "process_next_hop")

    def process_hop_io(self, in_hop, out_h=None, out_p=None, out_r=None):
        check(_lib.lib().zen_hpr_process_hop_io(self._h, _dptr(in_hop), _dptr(out_h) if out_h is not None else None,
                                                _dptr(out_p) if out_p is not None else None,
                                                _dptr(out_r) if out_r is not None else None), "process_hop_io")

    def synchronize(self):
        check(_lib.lib().zen_hpr_synchronize(self._h), "synchronize")

    def realtime_begin(self):
        """serve the following hops from the resident (persistent) kernel"""
        check(_lib.lib().zen_hpr_realtime_begin(self._h), "realtime_begin")

    def realtime_end(self):
        check(_lib.lib().zen_hpr_realtime_end(self._h), "realtime_end")

    def _state(self, which, n):
        self.synchronize()
        out = np.empty(n, dtype=np.float32)
        ptr = _lib.lib().zen_hpr_state_ptr(self._h, which)
        check(_lib.lib().zen_copy_to_host(out.ctypes.data, ptr, out.nbytes), "zen_copy_to_host")
        return out

    # the reference exposes these as public device_vectors (hps.h:182-197)
    @property
    def input(self):
        return self._state(0, self.nwin)

    @property
    def harmonic_out(self):
        return self._state(1, self.nwin)

    @property
    def percussive_out(self):
        return self._state(2, self.nwin)

    @property
    def residual_out(self):
        return self._state(3, self.nwin)

    def materialize(self):
        """dict of the reference's stft_width x nfft matrices, rebuilt from the ring state"""
        torch = _torch()
        n = self.stft_width * self.nfft
        names = ["sliding_stft", "s_mag", "harmonic_matrix", "percussive_matrix", "harmonic_mask", "percussive_mask",
                 "residual_mask"]
        bufs = [torch.zeros(2 * n if nm == "sliding_stft" else n, dtype=torch.float32, device="cuda") for nm in names]
        check(_lib.lib().zen_hpr_materialize(self._h, *[b.data_ptr() for b in bufs]), "zen_hpr_materialize")
        out = {}
        for nm, b in zip(names, bufs):
            a = b.cpu().numpy()
            out[nm] = a.view(np.complex64).reshape(self.stft_width, self.nfft) if nm == "sliding_stft" else a.reshape(
                self.stft_width, self.nfft)
        return out

    def run(self, audio, n_hops=None):
        """tests helper: feed host audio hop by hop, collect the first hop samples of each output"""
        torch = _torch()
        a = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).cuda()
        if n_hops is None:
            n_hops = a.numel() // self.hop
        outs = [torch.zeros(n_hops * self.hop, dtype=torch.float32, device="cuda") for _ in range(3)]
        hop = self.hop
        for i in range(n_hops):
            self.process_hop_io(a[i * hop:].data_ptr(), outs[0][i * hop:].data_ptr(), outs[1][i * hop:].data_ptr(),
                                outs[2][i * hop:].data_ptr())
        self.synchronize()
        return [o.cpu().numpy() for o in outs]


class HPRRealtime:
    """zen::hps::HPRRealtime<Backend::GPU> (libzen/libzen/hps.h:74-118, libzen/hps.cu:282-427)."""

    def __init__(self, fs, hop=256, beta=2.0, output_flags=OUTPUT_PERCUSSIVE, nocopybord=False):
        self.p_impl = HPR(fs, hop, beta, output_flags, MedianFilterDirection.TimeCausal, not nocopybord)

    def process_next_hop(self, in_hop):
        self.p_impl.process_next_hop(in_hop)

    def copy_harmonic(self, out_hop):
        check(_lib.lib().zen_hpr_copy_harmonic(self.p_impl._h, _dptr(out_hop)), "copy_harmonic")

    def copy_percussive(self, out_hop):
        check(_lib.lib().zen_hpr_copy_percussive(self.p_impl._h, _dptr(out_hop)), "copy_percussive")

    def copy_residual(self, out_hop):
        check(_lib.lib().zen_hpr_copy_residual(self.p_impl._h, _dptr(out_hop)), "copy_residual")

    def use_sse_filter(self):
        self.p_impl.use_sse_filter()

    def use_soft_mask(self):
        self.p_impl.use_soft_mask()

    def warmup(self, io, test_iters=1000):
        """hps.cu:392-409: 1000 hops of iota data, then reset_buffers()."""
        hop = self.p_impl.hop
        for i in range(test_iters):
            io.host_in[:hop] = np.arange(i * hop, (i + 1) * hop, dtype=np.float32)
            self.p_impl.process_next_hop(io.device_in)
            # process_next_hop only enqueues the hop kernel, which reads the mapped host_in later: wait for it before
            # host_in is refilled (in the reference the copy of in_hop has completed when the call returns)
            self.p_impl.synchronize()
        self.p_impl.reset_buffers()


class HPRIOffline:
    """zen::hps::HPRIOffline<Backend::GPU> (libzen/libzen/hps.h:29-72, libzen/hps.cu:21-221)."""

    def __init__(self, fs, hop_h=4096, hop_p=256, beta_h=2.0, beta_p=2.0, nocopybord=False):
        if hop_h % hop_p != 0:
            raise ZgException("hop_h and hop_p should be evenly divisible")
        self.fs, self.hop_h, self.hop_p, self.beta_h, self.beta_p = fs, hop_h, hop_p, beta_h, beta_p
        self.options = _lib.OPT_NOCOPYBORD if nocopybord else 0

    def use_sse_filter(self):
        self.options |= _lib.OPT_SSE

    def use_soft_mask(self):
        self.options |= _lib.OPT_SOFT_MASK

    def process(self, audio):
        a = np.ascontiguousarray(audio, dtype=np.float32)
        outs = [np.zeros(a.size, dtype=np.float32) for _ in range(3)]
        check(_lib.lib().zen_offline_process(self.fs, self.hop_h, self.hop_p, self.beta_h, self.beta_p, self.options,
                                             a.ctypes.data, a.size, *[o.ctypes.data for o in outs]), "zen_offline_process")
        return outs

    def process_device(self, audio_t):
        """device-resident variant: float32 CUDA tensor in, three CUDA tensors out"""
        torch = _torch()
        n = audio_t.numel()
        outs = [torch.empty(n, dtype=torch.float32, device=audio_t.device) for _ in range(3)]
        check(_lib.lib().zen_offline_process_device(self.fs, self.hop_h, self.hop_p, self.beta_h, self.beta_p, self.options,
                                                    audio_t.data_ptr(), n, *[o.data_ptr() for o in outs],
                                                    torch.cuda.current_stream().cuda_stream), "zen_offline_process_device")
        return outs


class HPRBatch:
    """Many independent HPRRealtime / HPR streams in one launch (zen_hpr_batch_*)."""

    def __init__(self, fs, hop, beta, output_flags, causal=True, nocopybord=False, sse=False, soft=False):
        opts = (_lib.OPT_NOCOPYBORD if nocopybord else 0) | (_lib.OPT_SSE if sse else 0) | (_lib.OPT_SOFT_MASK if soft else 0)
        b = ctypes.c_void_p()
        check(_lib.lib().zen_hpr_batch_create(ctypes.byref(b), fs, hop, beta, output_flags,
                                              _lib.TIME_CAUSAL if causal else _lib.TIME_ANTICAUSAL, opts, 1, 1),
              "zen_hpr_batch_create")
        self._b = b
        self.hop, self.flags = hop, output_flags

    def close(self):
        if getattr(self, "_b", None):
            _lib.lib().zen_hpr_batch_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def process(self, x, outs=None):
        """x: [n_streams, n_hops*hop] float32 CUDA tensor (rows contiguous, any even row stride).  Returns (H, P, R)
        tensors (None where disabled).  Equivalent to n_hops process_next_hop calls per stream: the residual of the
        soft-mask / SSE variants is all zeros (libzen/hps.cu:435-449, 562)."""
        torch = _torch()
        if x.dim() != 2 or x.dtype != torch.float32 or not x.is_cuda or (x.shape[1] > 1 and x.stride(1) != 1):
            raise ValueError("HPRBatch.process: x must be a 2-D float32 CUDA tensor with contiguous rows")
        n_streams, n = x.shape
        if n % self.hop != 0 or n == 0:
            raise ValueError("HPRBatch.process: the row length must be a positive multiple of hop")
        n_hops = n // self.hop
        if outs is None:
            outs = [torch.empty((n_streams, n), dtype=torch.float32, device=x.device) if self.flags & (1 << o) else None for o in range(3)]
        strides = set()
        for o in outs:
            if o is None:
                continue
            if o.shape != x.shape or o.dtype != torch.float32 or o.device != x.device or (n > 1 and o.stride(1) != 1):
                raise ValueError("HPRBatch.process: outputs must match x in shape / dtype / device and have contiguous rows")
            strides.add(o.stride(0) if n_streams > 1 else n)
        if len(strides) > 1:
            raise ValueError("HPRBatch.process: all outputs must share one row stride")
        out_stride = strides.pop() if strides else n
        in_stride = x.stride(0) if n_streams > 1 else n
        ptr = [o.data_ptr() if o is not None else None for o in outs]
        check(_lib.lib().zen_hpr_batch_process(self._b, x.data_ptr(), in_stride, n_streams, n_hops, ptr[0], ptr[1], ptr[2],
                                               out_stride, torch.cuda.current_stream().cuda_stream), "zen_hpr_batch_process")
        return outs

    @staticmethod
    def _host_rows(a, name, itemsize):
        """(address, row stride in elements) of a 2-D host array with contiguous rows"""
        if isinstance(a, np.ndarray):
            if a.ndim != 2 or a.itemsize != itemsize or (a.shape[1] > 1 and a.strides[1] != itemsize) or a.strides[0] % itemsize:
                raise ValueError("HPRBatch: %s must be 2-D with contiguous rows" % name)
            return a.ctypes.data, (a.strides[0] // itemsize if a.shape[0] > 1 else a.shape[1])
        if a.dim() != 2 or a.element_size() != itemsize or (a.shape[1] > 1 and a.stride(1) != 1):
            raise ValueError("HPRBatch: %s must be 2-D with contiguous rows" % name)
        return a.data_ptr(), (a.stride(0) if a.shape[0] > 1 else a.shape[1])

    def _process_host(self, fn, what, x, outs, itemsize, extra=()):
        n_streams, n = x.shape
        if n % self.hop != 0 or n == 0:
            raise ValueError("HPRBatch.%s: the row length must be a positive multiple of hop" % what)
        xa, xs = self._host_rows(x, "x", itemsize)
        oa, os_ = [None, None, None], set()
        for i, o in enumerate(outs):
            if o is None:
                continue
            if tuple(o.shape) != tuple(x.shape):
                raise ValueError("HPRBatch.%s: outputs must have the shape of x" % what)
            oa[i], st = self._host_rows(o, "outs[%d]" % i, itemsize)
            os_.add(st)
        if len(os_) > 1:
            raise ValueError("HPRBatch.%s: all outputs must share one row stride" % what)
        check(fn(self._b, xa, xs, n_streams, n // self.hop, oa[0], oa[1], oa[2], os_.pop() if os_ else n, *extra), what)

    def process_host(self, x, outs):
        """x, outs[*]: [n_streams, n_hops*hop] float32 host arrays (numpy or pinned torch tensors)."""
        self._process_host(_lib.lib().zen_hpr_batch_process_host, "zen_hpr_batch_process_host", x, outs, 4)

    def process_host_pcm16(self, x, outs, peaks=(None, None, None)):
        """The command line's sample format on both sides of the link (zen/offline.h:88-117, 180-223): x and outs[*] are
        [n_streams, n_hops*hop] int16 host arrays; every output comes back peak-normalised per stream and converted as
        libnyquist does.  peaks[*]: optional float32 arrays [n_streams] receiving the divisors."""
        pk = [None if p is None else p.ctypes.data for p in peaks]
        self._process_host(_lib.lib().zen_hpr_batch_process_host_pcm16, "zen_hpr_batch_process_host_pcm16", x, outs, 2, extra=pk)

    @property
    def last_kernel_ms(self):
        return float(_lib.lib().zen_hpr_batch_last_kernel_ms(self._b))

    @property
    def last_launches(self):
        return int(_lib.lib().zen_hpr_batch_last_launches(self._b))


def pcm16_decode_mono(pcm, channels=1):
    """batched libnyquist decode (Common.h:296-302, 669-675): int16 CUDA tensor [n_streams, n_frames * channels]
    (interleaved) -> float32 CUDA tensor [n_streams, n_frames]"""
    torch = _torch()
    assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.dim() == 2 and (pcm.shape[1] == 0 or pcm.stride(1) == 1)
    n_streams, n_frames = pcm.shape[0], pcm.shape[1] // channels
    out = torch.empty((n_streams, n_frames), dtype=torch.float32, device=pcm.device)
    if n_frames == 0:
        return out
    # (a one-row tensor may carry any stride for its first axis, 0 included)
    in_stride = pcm.stride(0) if n_streams > 1 else pcm.shape[1]
    check(_lib.lib().zen_pcm16_decode_mono(pcm.data_ptr(), in_stride, channels, n_streams, n_frames, out.data_ptr(), max(1, out.stride(0))),
          "zen_pcm16_decode_mono")
    return out


def pcm16_encode_normalized(x):
    """batched peak normalisation + PCM16 encode as the zen command line writes its outputs (zen/offline.h:180-192,
    libnyquist Common.cpp:332-337): float32 CUDA tensor [n_streams, n] -> (int16 [n_streams, n], float32 peaks [n_streams])"""
    torch = _torch()
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and (x.shape[1] == 0 or x.stride(1) == 1)
    out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    peaks = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
    if x.shape[1] == 0:
        return out, peaks
    in_stride = x.stride(0) if x.shape[0] > 1 else x.shape[1]
    check(_lib.lib().zen_pcm16_encode_normalized(x.data_ptr(), in_stride, x.shape[0], x.shape[1], out.data_ptr(), max(1, out.stride(0)),
                                                 peaks.data_ptr()), "zen_pcm16_encode_normalized")
    return out, peaks


class MPM:
    """McLeod pitch method of the reference's pitch-tracking demo (demos/pitch-tracking/pitch_detection.h:17-93,
    pitch.cpp:101-135), batched on the device: pitch(x) takes a float32 CUDA tensor [n_buffers, N] (or [N]) - e.g. the
    harmonic output of HPRBatch reshaped to hops - and returns the pitch of every buffer in Hz (-1 where the reference
    returns -1)."""

    def __init__(self, audio_buffer_size, sample_rate):
        if audio_buffer_size == 0:
            raise MemoryError("std::bad_alloc")          # pitch_detection.h:48-50
        self.N, self.sample_rate = int(audio_buffer_size), float(sample_rate)

    def pitch(self, x, want_nsdf=False):
        torch = _torch()
        one = x.dim() == 1
        x2 = x.reshape(1, -1) if one else x
        assert x2.is_cuda and x2.dtype == torch.float32 and x2.shape[1] == self.N and (self.N == 1 or x2.stride(1) == 1)
        nb = x2.shape[0]
        out = torch.empty(nb, dtype=torch.float32, device=x.device)
        nsdf = torch.empty((nb, self.N), dtype=torch.float32, device=x.device) if want_nsdf else None
        stride = x2.stride(0) if nb > 1 else self.N
        check(_lib.lib().zen_mpm_pitch(self.N, self.sample_rate, x2.data_ptr(), stride, nb, out.data_ptr(),
                                       nsdf.data_ptr() if want_nsdf else None, torch.cuda.current_stream().cuda_stream), "zen_mpm_pitch")
        res = out[0] if one else out
        return (res, nsdf[0] if one else nsdf) if want_nsdf else res


class OnsetDetectionFunction:
    """The onset detection function BTrack computes from the percussive output (demos/beat-tracking/OnsetDetection.cpp:60-131,
    complex spectral difference, half-wave rectified; FrameSize 512, HopSize 256), batched on the device:
    calculate_samples(x) takes a float32 CUDA tensor [n_streams, n_hops * 256] (or one stream, 1-D) - e.g. the percussive
    output of HPRBatch at hop 256 - and returns one sample per hop, [n_streams, n_hops]: the sequence the reference's
    calculate_sample() returns hop after hop on a fresh object."""
    FrameSize, HopSize = 512, 256

    def calculate_samples(self, x):
        torch = _torch()
        one = x.dim() == 1
        x2 = x.reshape(1, -1) if one else x
        assert x2.is_cuda and x2.dtype == torch.float32 and (x2.shape[1] == 0 or x2.stride(1) == 1)
        n_streams, n_hops = x2.shape[0], x2.shape[1] // self.HopSize
        out = torch.empty((n_streams, n_hops), dtype=torch.float32, device=x.device)
        if n_hops:
            stride = x2.stride(0) if n_streams > 1 else x2.shape[1]
            check(_lib.lib().zen_onset_csd(x2.data_ptr(), stride, n_streams, n_hops, out.data_ptr(), max(1, out.stride(0)),
                                           torch.cuda.current_stream().cuda_stream), "zen_onset_csd")
        return out[0] if one else out


class BTrack:
    """The beat tracker of the reference's beat-tracking demo (demos/beat-tracking/BTrack.cpp) behind
    OnsetDetectionFunction: host-side control logic that consumes one onset-detection sample per 256-sample hop.
    process_odf(samples) -> (beat_due [n] bool, tempo_bpm [n], cumulative_score [n]); process_percussive(x) runs the
    onset detection function on the device first (x: float32 CUDA tensor, one stream, 256 n_hops samples)."""

    def __init__(self, sample_rate):
        self._h = ctypes.c_void_p()
        check(_lib.lib().zen_btrack_create(ctypes.byref(self._h), int(sample_rate)), "zen_btrack_create")

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().zen_btrack_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def process_odf(self, samples):
        s = np.ascontiguousarray(samples, dtype=np.float32)
        n = s.size
        beat = np.zeros(n, np.uint8)
        tempo = np.zeros(n, np.float32)
        score = np.zeros(n, np.float32)
        check(_lib.lib().zen_btrack_process(self._h, s.ctypes.data, n, beat.ctypes.data, tempo.ctypes.data, score.ctypes.data), "zen_btrack_process")
        return beat.astype(bool), tempo, score

    def process_percussive(self, x):
        odf = OnsetDetectionFunction().calculate_samples(x.reshape(-1))
        return self.process_odf(odf.cpu().numpy())

    def tables(self):
        r = np.zeros(128, np.float32)
        t = np.zeros((41, 41), np.float32)
        check(_lib.lib().zen_btrack_tables(self._h, r.ctypes.data, t.ctypes.data), "zen_btrack_tables")
        return r, t
