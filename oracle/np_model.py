"""TEST INFRASTRUCTURE — NumPy twin of the window / border rules of the
reference's filter wrappers.  Slow, obviously-correct index arithmetic used to
cross-check oracle/hpr_oracle.c and the CUDA kernels on small matrices.

GPU geometry  : /root/reference/libzen/mfilt.h:93-199, 233-267 (NPP ROI/anchor
                arithmetic, nppiCopyWrapBorder); box.h:84-214.
CPU geometry  : /root/reference/libzen/mfilt.h:285-341 (IPP, centred window,
                ippBorderRepl); box.h:232-287.
Sizes         : /root/reference/libzen/hps.h:216-285.
"""
import numpy as np

CAUSAL, ANTICAUSAL, FREQUENCY = 0, 1, 2


class ZgException(ValueError):
    pass


def odd_len(filter_len: int) -> int:
    """mfilt.h:89-91: filter_len += 1 - filter_len % 2."""
    return filter_len + (1 - filter_len % 2)


def _check(T, F, filter_len, direction):
    # mfilt.h:80-87 / box.h:71-78: the check uses the length BEFORE it is made odd
    if (direction in (CAUSAL, ANTICAUSAL) and filter_len > T) or (direction == FREQUENCY and filter_len > F):
        raise ZgException("filter bigger than matrix dimension")


def window_indices_gpu(T, F, filter_len, direction, copy_bord):
    """Returns (axis, out_idx, taps) : for every written index i along `axis`
    the list of source indices that enter the window."""
    _check(T, F, filter_len, direction)
    L = odd_len(filter_len)
    mid = L // 2
    axis = 1 if direction == FREQUENCY else 0
    dim = F if axis == 1 else T
    out = {}
    if copy_bord:
        # wrap-padded source + full ROI: centred circular window in every direction
        for i in range(dim):
            out[i] = [(i - mid + k) % dim for k in range(L)]
    elif direction == CAUSAL:
        # anchor {0, L}: mask covers rows r-L .. r-1 (current row excluded)
        for r in range(L, T):
            out[r] = [r - L + k for k in range(L)]
    elif direction == ANTICAUSAL:
        for r in range(mid, mid + T - L):
            out[r] = [r - mid + k for k in range(L)]
    else:
        for c in range(0, F - L):
            out[c] = [c + k for k in range(L)]
    return axis, out


def median_filter_gpu(src, filter_len, direction, copy_bord, dst_init=None):
    src = np.asarray(src, dtype=np.float32)
    T, F = src.shape
    axis, taps = window_indices_gpu(T, F, filter_len, direction, copy_bord)
    dst = np.zeros_like(src) if dst_init is None else np.array(dst_init, dtype=np.float32, copy=True)
    L = odd_len(filter_len)
    for i, idx in taps.items():
        if axis == 0:
            dst[i, :] = np.sort(src[idx, :], axis=0)[L // 2, :]
        else:
            dst[:, i] = np.sort(src[:, idx], axis=1)[:, L // 2]
    return dst


def box_filter_gpu(src, filter_len, direction, dst_init=None):
    """BoxFilterGPU always wrap-pads (box.h:194-213): circular centred mean."""
    src = np.asarray(src, dtype=np.float32)
    T, F = src.shape
    axis, taps = window_indices_gpu(T, F, filter_len, direction, True)
    dst = np.zeros_like(src) if dst_init is None else np.array(dst_init, dtype=np.float32, copy=True)
    L = odd_len(filter_len)
    with np.errstate(all="ignore"):
        for i, idx in taps.items():
            if axis == 0:
                dst[i, :] = (src[idx, :].astype(np.float64).sum(axis=0) / L).astype(np.float32)
            else:
                dst[:, i] = (src[:, idx].astype(np.float64).sum(axis=1) / L).astype(np.float32)
    return dst


def _taps_cpu(dim, L):
    mid = L // 2
    return {i: [min(max(i - mid + k, 0), dim - 1) for k in range(L)] for i in range(dim)}


def median_filter_cpu(src, filter_len, direction):
    src = np.asarray(src, dtype=np.float32)
    T, F = src.shape
    _check(T, F, filter_len, direction)
    L = odd_len(filter_len)
    dst = np.zeros_like(src)
    if direction == FREQUENCY:
        for i, idx in _taps_cpu(F, L).items():
            dst[:, i] = np.sort(src[:, idx], axis=1)[:, L // 2]
    else:
        for i, idx in _taps_cpu(T, L).items():
            dst[i, :] = np.sort(src[idx, :], axis=0)[L // 2, :]
    return dst


def box_filter_cpu(src, filter_len, direction):
    src = np.asarray(src, dtype=np.float32)
    T, F = src.shape
    _check(T, F, filter_len, direction)
    L = odd_len(filter_len)
    dst = np.zeros_like(src)
    with np.errstate(all="ignore"):
        if direction == FREQUENCY:
            for i, idx in _taps_cpu(F, L).items():
                dst[:, i] = (src[:, idx].astype(np.float64).sum(axis=1) / L).astype(np.float32)
        else:
            for i, idx in _taps_cpu(T, L).items():
                dst[i, :] = (src[idx, :].astype(np.float64).sum(axis=0) / L).astype(np.float32)
    return dst


def hpr_geometry(fs: float, hop: int, causal: bool):
    """hps.h:222-230, 265-268 evaluated with the same float/double mix:
    l_harm = roundf(0.2 / ((float)(nfft - hop) / fs))  -> float divide, double divide
    l_perc = roundf(500 / (fs / (float)nfft))."""
    fs32 = np.float32(fs)
    nwin, nfft = 2 * hop, 4 * hop
    denom = np.float32(np.float32(nfft - hop) / fs32)
    l_harm = int(np.round(np.float32(0.2 / float(denom))))  # roundf of a positive value
    l_perc = int(np.floor(np.float32(500.0) / np.float32(fs32 / np.float32(nfft)) + np.float32(0.5)))
    lag = 1 if causal else l_harm
    return dict(nwin=nwin, nfft=nfft, l_harm=l_harm, l_perc=l_perc, lag=lag, stft_width=2 * l_harm)


# ---- the on-disk format either side of the path ------------------------------
# /root/reference/vendor/libnyquist/include/libnyquist/Common.h:296-302 (int16_to_float32, float32_to_int16),
# Common.h:669-675 (StereoToMono), vendor/libnyquist/src/Common.cpp:332-337 (lroundf, no dither),
# /root/reference/zen/offline.h:104-117, 180-192 and zen/fakert.h:117-130, 259-268 (mono fold, peak normalisation).

def pcm16_decode_mono(pcm, channels):
    """int16 [n_frames * channels] (interleaved) -> float32 [n_frames]: (float)s / 32767.f, stereo as (l + r) / 2.0f"""
    f = pcm.astype(np.float32) / np.float32(32767.0)
    if channels == 2:
        f = (f[0::2] + f[1::2]) / np.float32(2.0)
    return f.astype(np.float32)


def lroundf(v):
    """C lroundf on float32 values: nearest integer, halfway cases away from zero"""
    v64 = v.astype(np.float64)
    return np.trunc(v64 + np.copysign(0.5, v64)).astype(np.int64)


def pcm16_encode_normalized(x):
    """float32 [n] -> (int16 [n], peak): x / max(-min, max), then (int16_t)lroundf(x * 32767.f).
    A silent signal (peak 0) stays silent: the reference divides 0 by 0 and hands NaN to lroundf, which its x86-64
    build turns into PCM16 zeros as well (tests/golden/nyq_pcm.npz: silence_out, from the compiled libnyquist).
    Pinned by tests/test_pcm.py::test_pcm16_oracle_pinned_by_libnyquist."""
    x = x.astype(np.float32)
    if x.size == 0:
        return np.zeros(0, np.int16), np.float32(0.0)
    peak = np.float32(max(-x.min(), x.max()))
    if not peak > 0:
        return np.zeros(x.size, np.int16), np.float32(0.0)
    y = (x / peak).astype(np.float32) * np.float32(32767.0)
    return lroundf(y.astype(np.float32)).astype(np.int16), peak
