"""Pins the oracle (oracle/hpr_oracle.c) — CPU only:
  * against the reference's own known-answer tests (libzen/mfilt.test.cu,
    libzen/box.test.cu, libzen/fftw.test.cu, restated here);
  * against golden vectors produced by the UNMODIFIED reference compiled from
    /root/reference and run on a B200 (tests/golden, oracle/ref/probe_ref_gpu.py);
  * against the NumPy twin of the window rules (oracle/np_model.py).
"""
import hashlib
import os
import re

import numpy as np
import pytest

from oracle import np_model
from tests.golden_inputs import median_case_input
from tests.util import peak_norm_err
from zen_b200.synth import synth_audio

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SENT = np.float32(-777.0)


# ------------------------------------------------ reference known-answer tests ---

def cross(T, F):
    """libzen/mfilt.test.cu:32-38: zeros with row T/2 = 5 and column F/2 = 8"""
    m = np.zeros((T, F), dtype=np.float32)
    m[T // 2, :] = 5
    m[:, F // 2] = 8
    return m


@pytest.mark.parametrize("T,F,L", [(9, 9, 3), (10, 20, 5), (1024, 17, 5)])
def test_mfilt_test_cu_gpu_nocopybord(oracle, T, F, L):
    """mfilt.test.cu:117-301: causal -> column stays 8 for i > L; anticausal -> 8 for 2 < i < T-3;
    frequency -> row stays 5 for j < F-L; elsewhere 0"""
    src = cross(T, F)
    out = oracle.median_filter(oracle.GEOM_GPU, src, L, oracle.CAUSAL, False)
    for i in range(T):
        for j in range(F):
            if j == F // 2 and i > L:
                assert out[i, j] == 8
            elif j != F // 2:
                assert out[i, j] == 0
    out = oracle.median_filter(oracle.GEOM_GPU, src, L, oracle.ANTICAUSAL, False)
    for i in range(T):
        for j in range(F):
            if j == F // 2 and 2 < i < T - 3:
                assert out[i, j] == 8
            elif j != F // 2:
                assert out[i, j] == 0
    out = oracle.median_filter(oracle.GEOM_GPU, src, L, oracle.FREQUENCY, False)
    for i in range(T):
        for j in range(F):
            if i == T // 2 and j < F - L:
                assert out[i, j] == 5
            elif i != T // 2:
                assert out[i, j] == 0


@pytest.mark.parametrize("T,F,L", [(9, 9, 3), (10, 20, 5), (1024, 17, 5)])
def test_mfilt_test_cu_gpu_copybord(oracle, T, F, L):
    """mfilt.test.cu:701-886: with copy-border the column / the row survives everywhere"""
    src = cross(T, F)
    for d in (oracle.CAUSAL, oracle.ANTICAUSAL):
        out = oracle.median_filter(oracle.GEOM_GPU, src, L, d, True)
        exp = np.zeros_like(src)
        exp[:, F // 2] = 8
        assert np.array_equal(out, exp)
    out = oracle.median_filter(oracle.GEOM_GPU, src, L, oracle.FREQUENCY, True)
    exp = np.zeros_like(src)
    exp[T // 2, :] = 5
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("T,F,L", [(9, 9, 3), (10, 20, 5), (1024, 128, 5)])
def test_mfilt_test_cu_cpu(oracle, T, F, L):
    """mfilt.test.cu:407-591 (IPP path): the same conditional expectations as GPU-nocopybord"""
    src = cross(T, F)
    out = oracle.median_filter(oracle.GEOM_CPU, src, L, oracle.CAUSAL, False)
    for i in range(T):
        if i > L:
            assert out[i, F // 2] == 8
    out = oracle.median_filter(oracle.GEOM_CPU, src, L, oracle.FREQUENCY, False)
    for j in range(F - L):
        assert out[T // 2, j] == 5


@pytest.mark.parametrize("geom", [0, 1])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_filter_len_bigger_than_dim_throws(oracle, geom, d):
    """mfilt.test.cu:235-244, 525-534, 818-829; box.h:71-78"""
    src = np.zeros((9, 9), np.float32)
    with pytest.raises(ValueError):
        oracle.median_filter(geom, src, 10, d, True)
    with pytest.raises(ValueError):
        oracle.box_filter(geom, src, 10, d)


@pytest.mark.parametrize("L,expect", [(3, 32.0), (5, 48.0)])
def test_box_test_cu_values(oracle, L, expect):
    """box.test.cu:124-201: reciprocal -> box(time) -> reciprocal*(L+1) on a column of 8 gives 8*(L+1):
    proves the box filter is a MEAN and the (l+1) convention of hps.cu:599-604"""
    T, F = 10, 20
    src = np.full((T, F), 8.0, dtype=np.float32)
    rec = (np.float32(1.0) / src).astype(np.float32)
    for d in (oracle.CAUSAL, oracle.ANTICAUSAL):
        out = oracle.box_filter(oracle.GEOM_GPU, rec, L, d)
        res = (np.float32(1.0) / out) * np.float32(L + 1.0)
        assert np.allclose(res, expect, rtol=1e-6)


@pytest.mark.parametrize("n", [64, 1024, 16384])
def test_fftw_test_cu_tolerance(oracle, n):
    """fftw.test.cu:83-101: forward and backward agree with a second implementation within 2e-4, unnormalised"""
    rng = np.random.default_rng(n)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    assert np.abs(oracle.fft(x) - np.fft.fft(x.astype(np.complex128))).max() <= 2e-4
    assert np.abs(oracle.fft(x, inverse=True) - np.fft.ifft(x.astype(np.complex128)) * n).max() <= 2e-4
    back = oracle.fft(oracle.fft(x), inverse=True) / n
    assert np.abs(back - x).max() <= 1e-5


# ------------------------------------------------------ golden: NPP / cuFFT ---

def test_oracle_median_equals_npp_golden(oracle):
    """90 probe cases on a B200: sentinel-prefilled dst, 3 directions x 2 border modes, ties, ZgException cases"""
    d = np.load(os.path.join(GOLD, "npp_median.npz"))
    keys = sorted(set(re.match(r"(T\d+_F\d+_L\d+_d\d_cb\d)", k).group(1) for k in d.files))
    assert len(keys) == 90
    for k in keys:
        T, F, L, dr, cb = map(int, re.match(r"T(\d+)_F(\d+)_L(\d+)_d(\d)_cb(\d)", k).groups())
        if k + "_zgexception" in d.files:
            with pytest.raises(ValueError):
                oracle.median_filter(oracle.GEOM_GPU, np.zeros((T, F), np.float32), L, dr, cb)
            continue
        src = d[k + "_src"] if k + "_src" in d.files else median_case_input(int(d[k + "_seed"]), T, F)
        out = oracle.median_filter(oracle.GEOM_GPU, src, L, dr, cb, dst_init=np.full((T, F), SENT, np.float32))
        if k + "_dst" in d.files:
            assert np.array_equal(out.view(np.uint32), d[k + "_dst"].view(np.uint32)), k
            twin = np_model.median_filter_gpu(src, L, dr, cb, dst_init=np.full((T, F), SENT, np.float32))
            assert np.array_equal(twin.view(np.uint32), out.view(np.uint32)), k
        else:
            assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest() == d[k + "_dstsha"].tobytes(), k


def test_oracle_box_vs_npp_golden(oracle):
    d = np.load(os.path.join(GOLD, "npp_box.npz"))
    keys = sorted(set(re.match(r"(T\d+_F\d+_L\d+_d\d_[a-z]+)", k).group(1) for k in d.files))
    for k in keys:
        T, F, L, dr = map(int, re.match(r"T(\d+)_F(\d+)_L(\d+)_d(\d)", k).groups())
        if k + "_zgexception" in d.files:
            with pytest.raises(ValueError):
                oracle.box_filter(oracle.GEOM_GPU, np.zeros((T, F), np.float32), L, dr)
            continue
        with np.errstate(all="ignore"):
            out = oracle.box_filter(oracle.GEOM_GPU, d[k + "_src"], L, dr)
        ref = d[k + "_dst"]
        assert np.array_equal(np.isinf(out), np.isinf(ref)) and not np.isnan(ref).any()
        fin = np.isfinite(ref)
        assert (np.abs(out[fin] - ref[fin]) / np.abs(ref[fin])).max() <= 1e-6, k


@pytest.mark.parametrize("n", [64, 1024, 4096])
def test_oracle_fft_vs_cufft_golden(oracle, n):
    d = np.load(os.path.join(GOLD, "cufft.npz"))
    x = d["n%d_x" % n]
    assert np.abs(oracle.fft(x) - d["n%d_fwd" % n]).max() <= 2e-4
    assert np.abs(oracle.fft(x, inverse=True) - d["n%d_inv" % n]).max() <= 2e-4


# ------------------------------------------- golden: reference separated audio ---

@pytest.mark.parametrize("name", ["rt4096_cb", "rt2048_nocb", "ac4096_cb", "ac4096_nocb"])
def test_oracle_vs_reference_gpu_audio(oracle, name):
    """deterministic configs of the reference GPU path (stft_width == 2): 1e-4 / 80 dB after peak normalisation"""
    d = np.load(os.path.join(GOLD, "ref_gpu_audio_%s.npz" % name))
    be, fs, hop, beta, flags, caus, cb, sse, soft, n_hops, seed = d["params"]
    hop, n_hops = int(hop), int(n_hops)
    audio = synth_audio(n_hops * hop, seed=int(seed), fs=int(fs))
    assert hashlib.sha256(audio.tobytes()).digest() == d["audio_sha"].tobytes()
    o = oracle.OracleHPR(oracle.GEOM_GPU, float(fs), hop, float(beta), int(flags), int(caus), bool(cb))
    nwin, nfft, l_harm, l_perc, lag, W, cola = d["geom"]
    assert (o.nwin, o.nfft, o.l_harm, o.l_perc, o.lag, o.stft_width) == tuple(int(v) for v in (nwin, nfft, l_harm, l_perc, lag, W))
    assert np.float32(o.cola) == np.float32(cola)
    outs = o.run(audio)
    for nm, a in zip(("harmonic", "percussive", "residual"), outs):
        err, snr = peak_norm_err(a, d[nm])
        assert err <= 1e-4 and (snr >= 80.0 or np.all(d[nm] == 0)), (nm, err, snr)


@pytest.mark.parametrize("name", ["rt1024_cb", "rt1024_nocb", "rt256_48k_cb", "ac256_48k_nocb", "rt512_soft"])
def test_oracle_stages_on_reference_dumps(oracle, name):
    """self-consistent final-hop dumps of the reference GPU path: its s_mag -> our filters == its matrices (bit-exact),
    its matrices -> our mask formulas == its masks (exact for the hard mask)"""
    d = np.load(os.path.join(GOLD, "ref_gpu_stages_%s.npz" % name))
    be, fs, hop, beta, flags, caus, cb, sse, soft, n_hops, seed = d["params"]
    nwin, nfft, l_harm, l_perc, lag, W = (int(v) for v in d["geom"][:6])
    s_mag = d["final_s_mag"]
    assert np.array_equal(s_mag[W - lag], np.abs(d["final_stft_row"]).astype(np.float32)) or \
        np.abs(s_mag[W - lag] - np.hypot(d["final_stft_row"].real, d["final_stft_row"].imag)).max() <= 4e-6 * s_mag.max()
    H = oracle.median_filter(oracle.GEOM_GPU, s_mag, l_harm, int(caus), bool(cb))
    P = oracle.median_filter(oracle.GEOM_GPU, s_mag, l_perc, oracle.FREQUENCY, bool(cb))
    assert np.array_equal(H.view(np.uint32), d["final_harmonic_matrix"].view(np.uint32))
    assert np.array_equal(P.view(np.uint32), d["final_percussive_matrix"].view(np.uint32))
    r = W - lag
    eps = np.float32(np.finfo(np.float32).eps)
    Hr, Pr = H[r], P[r]
    if not soft:
        with np.errstate(all="ignore"):
            mp = ((Pr / (Hr + eps)) >= np.float32(beta)).astype(np.float32)
            mh = ((Hr / (Pr + eps)) >= np.float32(np.float32(beta) - eps)).astype(np.float32)
        assert np.array_equal(mp, d["final_percussive_mask"][r])
        assert np.array_equal(mh, d["final_harmonic_mask"][r])
        assert np.array_equal((1 - (mh + mp)).astype(np.float32), d["final_residual_mask"][r])
    else:
        p = np.float32(int(beta))
        with np.errstate(all="ignore"):
            mp = np.power(Pr, p) / (np.power(Pr, p) + np.power(Hr, p) + eps)
        assert np.nanmax(np.abs(mp - d["final_percussive_mask"][r])) <= 2e-6


def test_oracle_sse_stage_on_reference_dump(oracle):
    d = np.load(os.path.join(GOLD, "ref_gpu_stages_rt512_sse.npz"))
    nwin, nfft, l_harm, l_perc, lag, W = (int(v) for v in d["geom"][:6])
    rec = d["final_reciprocal"]
    with np.errstate(all="ignore"):
        Hm = oracle.box_filter(oracle.GEOM_GPU, rec, l_harm, oracle.CAUSAL)
        Pm = oracle.box_filter(oracle.GEOM_GPU, rec, l_perc, oracle.FREQUENCY)
        H = (np.float32(1.0) / Hm) * np.float32(l_harm + 1.0)
        P = (np.float32(1.0) / Pm) * np.float32(l_perc + 1.0)
    for mine, ref in ((H, d["final_harmonic_matrix"]), (P, d["final_percussive_matrix"])):
        fin = np.isfinite(ref) & (ref > 0)
        assert np.array_equal(np.isfinite(mine), np.isfinite(ref))
        assert (np.abs(mine[fin] - ref[fin]) / ref[fin]).max() <= 2e-6


def test_oracle_offline_vs_reference(oracle):
    """HPRIOffline: GPU path pass 1 (deterministic) and the CPU path, both from the unmodified reference on the box"""
    audio = synth_audio(10 * 4096 + 11, seed=21)
    d = np.load(os.path.join(GOLD, "ref_offline_gpu.npz"))
    outs = oracle.offline_process(oracle.GEOM_GPU, 44100.0, 4096, 256, 2.5, 2.5, audio)
    err, snr = peak_norm_err(outs[0], d["harmonic"])
    assert err <= 1e-4 and snr >= 80.0
    assert int(d["residual_all_zero"]) == 1 and np.all(outs[2] == 0)
    d = np.load(os.path.join(GOLD, "ref_offline_gpu_nocb.npz"))
    outs = oracle.offline_process(oracle.GEOM_GPU, 44100.0, 4096, 256, 2.5, 2.5, audio, nocopybord=True)
    assert np.all(d["harmonic"] == 0) and np.all(outs[0] == 0)      # degenerate H = 0 case of SURVEY.md 8(a')
    d = np.load(os.path.join(GOLD, "ref_offline_cpu.npz"))
    outs = oracle.offline_process(oracle.GEOM_CPU, 44100.0, 4096, 256, 2.5, 2.5, audio)
    assert np.array_equal(outs[0], d["percussive"]) and np.array_equal(outs[1], d["percussive"]) and np.array_equal(outs[2], outs[0])
    assert np.array_equal(d["harmonic"], d["percussive"])           # CPU path returns percussive three times (hps.cu:278-279)


# --------------------------------------------------------------- host logic ---

@pytest.mark.parametrize("fs,hop,expect", [
    (44100.0, 256, (11, 12, 22)), (44100.0, 512, (6, 23, 12)), (44100.0, 1024, (3, 46, 6)), (44100.0, 2048, (1, 93, 2)),
    (44100.0, 4096, (1, 186, 2)), (48000.0, 256, (12, 11, 24))])
def test_geometry_table(oracle, fs, hop, expect):
    """SURVEY.md section 8 size table (hps.h:227-230 float/double mix)"""
    g = oracle.geometry(fs, hop, True)
    assert (g.l_harm, g.l_perc, g.stft_width) == expect and g.lag == 1
    assert oracle.geometry(fs, hop, False).lag == expect[0]
    m = np_model.hpr_geometry(fs, hop, True)
    assert (m["l_harm"], m["l_perc"], m["stft_width"]) == expect


def test_fakert_chunks(oracle):
    """zen/fakert.h:15-34: the trailing chunk is never produced"""
    assert oracle.fakert_n_chunks(161571, 1024) == 157
    assert oracle.fakert_n_chunks(161571, 256) == 631     # README.md:129 "631 chunks"
    assert oracle.fakert_n_chunks(1024, 1024) == 0
    assert oracle.fakert_n_chunks(2048, 1024) == 1


def test_oracle_cpu_vs_gpu_geometry_differ_only_at_borders(oracle):
    """hps.test.cu:230-284: GPU copybord != nocopybord; CPU ignores the flag"""
    audio = synth_audio(60 * 256, seed=3, fs=48000)
    runs = {}
    for geom in (0, 1):
        for cb in (True, False):
            runs[(geom, cb)] = oracle.OracleHPR(geom, 48000.0, 256, 2.0, 7, oracle.CAUSAL, cb).run(audio)
    assert any(not np.array_equal(a, b) for a, b in zip(runs[(0, True)], runs[(0, False)]))
    assert all(np.array_equal(a, b) for a, b in zip(runs[(1, True)], runs[(1, False)]))
