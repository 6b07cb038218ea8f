"""ctypes binding of zen_b200/lib/libzen_b200.so (the C ABI in include/zen_b200.h).

There is no fallback: if the library is missing, or a call reports an error,
an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZEN_B200_LIB") or os.path.join(_HERE, "lib", "libzen_b200.so")  # override: A/B builds only

ZEN_OK, ZEN_ERR_GEOMETRY, ZEN_ERR_CUDA, ZEN_ERR_UNSUPPORTED, ZEN_ERR_ARG = 0, -1, -2, -3, -4
TIME_CAUSAL, TIME_ANTICAUSAL, FREQUENCY = 0, 1, 2
OUTPUT_HARMONIC, OUTPUT_PERCUSSIVE, OUTPUT_RESIDUAL = 1, 2, 4
WIN_SQRT_VON_HANN, WIN_VON_HANN = 0, 1
OPT_SSE, OPT_SOFT_MASK, OPT_NOCOPYBORD = 1, 2, 4


class ZgException(RuntimeError):
    """zen::ZgException (libzen/libzen/zen.h:8-12)."""


class ZenCudaError(RuntimeError):
    pass


class ZenGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("hop", "nwin", "nfft", "l_harm", "l_perc", "lag", "stft_width")] + [
        ("cola_factor", ctypes.c_float)]


class ZenIO(ctypes.Structure):
    _fields_ = [("host_in", ctypes.c_void_p), ("host_out", ctypes.c_void_p), ("device_in", ctypes.c_void_p),
                ("device_out", ctypes.c_void_p), ("size", ctypes.c_size_t)]


EXPORTS = [
    "zen_b200_version", "zen_device_count", "zen_hpr_geometry", "zen_window", "zen_io_alloc", "zen_io_free",
    "zen_median_filter", "zen_box_filter", "zen_fft_c2c", "zen_hpr_create", "zen_hpr_destroy",
    "zen_hpr_use_sse_filter", "zen_hpr_use_soft_mask", "zen_hpr_reset_buffers", "zen_hpr_get_geometry",
    "zen_hpr_process_next_hop", "zen_hpr_copy_harmonic", "zen_hpr_copy_percussive", "zen_hpr_copy_residual",
    "zen_hpr_process_hop_io", "zen_hpr_synchronize", "zen_hpr_state_ptr", "zen_hpr_materialize",
    "zen_hpr_batch_create", "zen_hpr_batch_destroy", "zen_hpr_batch_process", "zen_hpr_batch_process_host",
    "zen_hpr_batch_last_launches", "zen_hpr_batch_last_kernel_ms", "zen_offline_process",
    "zen_offline_process_device", "zen_copy_to_host", "zen_copy_to_device", "zen_fakert_run", "zen_host_alloc", "zen_host_free", "zen_hpr_bind_state", "zen_hpr_realtime_begin", "zen_hpr_realtime_end", "zen_hpr_realtime_stamps",
    "zen_rt_pack_groups", "zen_rt_unpack_groups", "zen_pcm16_decode_mono", "zen_pcm16_encode_normalized",
    "zen_rt_split_ranges", "zen_pcm16_decode_mono_async", "zen_pcm16_peaks_async", "zen_pcm16_encode_with_peaks_async",
    "zen_pcm16_encode_normalized_async", "zen_hpr_batch_process_host_pcm16", "zen_hpr_wait_input_consumed", "zen_hpr_realtime_stamps_rank1", "zen_mpm_pitch", "zen_onset_csd", "zen_btrack_create", "zen_btrack_destroy", "zen_btrack_process", "zen_btrack_tables",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("zen_b200: %s is missing; build it with `make -C zen_b200/csrc` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cf, cu, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint, ctypes.c_long
    L.zen_b200_version.restype = ctypes.c_char_p
    L.zen_hpr_geometry.argtypes = [cf, ci, ci, ctypes.POINTER(ZenGeometry)]
    L.zen_window.argtypes = [ci, ci, vp]
    L.zen_io_alloc.argtypes = [ctypes.POINTER(ZenIO), ctypes.c_size_t]
    L.zen_io_free.argtypes = [ctypes.POINTER(ZenIO)]
    L.zen_io_free.restype = None
    L.zen_pcm16_decode_mono.argtypes = [vp, cl, ci, ci, cl, vp, cl]
    L.zen_pcm16_encode_normalized.argtypes = [vp, cl, ci, cl, vp, cl, vp]
    L.zen_pcm16_decode_mono_async.argtypes = [vp, cl, ci, ci, cl, vp, cl, vp]
    L.zen_pcm16_peaks_async.argtypes = [vp, cl, ci, cl, vp, vp]
    L.zen_pcm16_encode_with_peaks_async.argtypes = [vp, cl, ci, cl, vp, vp, cl, vp]
    L.zen_pcm16_encode_normalized_async.argtypes = [vp, cl, ci, cl, vp, cl, vp, vp]
    L.zen_hpr_batch_process_host_pcm16.argtypes = [vp, vp, cl, ci, cl, vp, vp, vp, cl, vp, vp, vp]
    L.zen_mpm_pitch.argtypes = [ci, cf, vp, cl, ci, vp, vp, vp]
    L.zen_onset_csd.argtypes = [vp, cl, ci, cl, vp, cl, vp]
    L.zen_btrack_create.argtypes = [ctypes.POINTER(vp), ci]
    L.zen_btrack_destroy.argtypes = [vp]
    L.zen_btrack_destroy.restype = None
    L.zen_btrack_process.argtypes = [vp, vp, cl, vp, vp, vp]
    L.zen_btrack_tables.argtypes = [vp, vp, vp]
    L.zen_median_filter.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp]
    L.zen_box_filter.argtypes = [ci, ci, ci, ci, vp, vp, vp]
    L.zen_fft_c2c.argtypes = [ci, vp, ci, vp]
    L.zen_hpr_create.argtypes = [ctypes.POINTER(vp), cf, ci, cf, cu, ci, ci]
    L.zen_hpr_destroy.argtypes = [vp]
    L.zen_hpr_destroy.restype = None
    for n in ("zen_hpr_use_sse_filter", "zen_hpr_use_soft_mask", "zen_hpr_reset_buffers", "zen_hpr_synchronize", "zen_hpr_wait_input_consumed"):
        getattr(L, n).argtypes = [vp]
    L.zen_hpr_get_geometry.argtypes = [vp, ctypes.POINTER(ZenGeometry)]
    L.zen_hpr_process_next_hop.argtypes = [vp, vp]
    for n in ("zen_hpr_copy_harmonic", "zen_hpr_copy_percussive", "zen_hpr_copy_residual"):
        getattr(L, n).argtypes = [vp, vp]
    L.zen_hpr_process_hop_io.argtypes = [vp, vp, vp, vp, vp]
    L.zen_hpr_state_ptr.argtypes = [vp, ci]
    L.zen_hpr_state_ptr.restype = vp
    L.zen_hpr_materialize.argtypes = [vp] * 8
    L.zen_hpr_batch_create.argtypes = [ctypes.POINTER(vp), cf, ci, cf, cu, ci, ci, ci, cl]
    L.zen_hpr_batch_destroy.argtypes = [vp]
    L.zen_hpr_batch_destroy.restype = None
    L.zen_hpr_batch_process.argtypes = [vp, vp, cl, ci, cl, vp, vp, vp, cl, vp]
    L.zen_hpr_batch_process_host.argtypes = [vp, vp, cl, ci, cl, vp, vp, vp, cl]
    L.zen_hpr_batch_last_launches.argtypes = [vp]
    L.zen_hpr_batch_last_launches.restype = cl
    L.zen_hpr_batch_last_kernel_ms.argtypes = [vp]
    L.zen_hpr_batch_last_kernel_ms.restype = cf
    L.zen_offline_process.argtypes = [cf, ci, ci, cf, cf, ci, vp, cl, vp, vp, vp]
    L.zen_offline_process_device.argtypes = [cf, ci, ci, cf, cf, ci, vp, cl, vp, vp, vp, vp]
    L.zen_fakert_run.argtypes = [cf, ci, cf, ci, vp, cl, ci, ci, vp, vp]
    L.zen_hpr_realtime_begin.argtypes = [vp]
    L.zen_hpr_realtime_end.argtypes = [vp]
    L.zen_hpr_realtime_stamps.argtypes = [vp, vp]
    L.zen_hpr_realtime_stamps_rank1.argtypes = [vp, vp]
    L.zen_hpr_bind_state.argtypes = [vp, vp, vp, vp, vp]
    L.zen_host_alloc.argtypes = [ctypes.c_size_t]
    L.zen_host_alloc.restype = vp
    L.zen_host_free.argtypes = [vp]
    L.zen_host_free.restype = None
    L.zen_copy_to_host.argtypes = [vp, vp, ctypes.c_size_t]
    L.zen_copy_to_device.argtypes = [vp, vp, ctypes.c_size_t]
    _lib = L
    return L


def check(rc, what=""):
    if rc == ZEN_OK:
        return
    if rc == ZEN_ERR_GEOMETRY:
        raise ZgException(what or "invalid geometry")
    if rc == ZEN_ERR_CUDA:
        raise ZenCudaError("zen_b200: CUDA failure in %s (no CPU fallback exists)" % what)
    if rc == ZEN_ERR_UNSUPPORTED:
        raise NotImplementedError("zen_b200: unsupported size in %s" % what)
    raise ValueError("zen_b200: bad argument to %s" % what)
