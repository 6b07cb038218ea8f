"""Parity of the CUDA path (through the C ABI) against the oracle and against
golden vectors captured from the unmodified reference on a B200.

Bars (BASELINE.json north_star): median filter bit-exact; separated audio
max-abs <= 1e-4 and SNR >= 80 dB on peak-normalised output.
"""
import glob
import hashlib
import os
import re

import numpy as np
import pytest

from tests.golden_inputs import median_case_input
from tests.util import flip_aware_compare, peak_norm_err
from zen_b200.synth import synth_audio

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SENT = np.float32(-777.0)
TOL_ABS, TOL_SNR = 1e-4, 80.0


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def zen():
    from zen_b200 import hps
    return hps


# ------------------------------------------------------------------ median ---

def _median_cases():
    d = np.load(os.path.join(GOLD, "npp_median.npz"))
    keys = sorted(set(re.match(r"(T\d+_F\d+_L\d+_d\d_cb\d)", k).group(1) for k in d.files))
    return d, keys


def test_median_bit_exact_vs_npp(torch, zen):
    """MedianFilterGPU::filter == nppiFilterMedian_32f_C1R as driven by libzen/mfilt.h, incl. untouched cells."""
    d, keys = _median_cases()
    assert len(keys) == 90
    n_checked = 0
    for k in keys:
        T, F, L, dr, cb = map(int, re.match(r"T(\d+)_F(\d+)_L(\d+)_d(\d)_cb(\d)", k).groups())
        if k + "_zgexception" in d.files:
            with pytest.raises(zen.ZgException):
                zen.MedianFilterGPU(T, F, L, dr, bool(cb))
            continue
        src = d[k + "_src"] if k + "_src" in d.files else median_case_input(int(d[k + "_seed"]), T, F)
        s = torch.from_numpy(src).cuda()
        o = torch.full((T, F), float(SENT), dtype=torch.float32, device="cuda")
        zen.MedianFilterGPU(T, F, L, dr, bool(cb)).filter(s, o)
        got = o.cpu().numpy()
        if k + "_dst" in d.files:
            assert np.array_equal(got.view(np.uint32), d[k + "_dst"].view(np.uint32)), k
        else:
            sha = hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest()
            assert sha == d[k + "_dstsha"].tobytes(), k
        n_checked += 1
    assert n_checked >= 60


@pytest.mark.parametrize("T,F,L,dr,cb", [(64, 300, 47, 2, 1), (5, 5000, 187, 2, 1), (5, 5000, 93, 2, 0), (300, 70, 21, 0, 1),
                                          (300, 70, 21, 1, 0), (40, 33, 33, 2, 1), (1000, 1000, 11, 0, 0), (1000, 1000, 11, 2, 1),
                                          (1, 64, 5, 2, 1), (7, 1, 1, 0, 1)])
def test_median_bit_exact_vs_oracle(torch, zen, oracle, T, F, L, dr, cb):
    rng = np.random.default_rng(T * 7 + F + L)
    src = rng.standard_normal((T, F)).astype(np.float32)
    src[rng.random((T, F)) < 0.05] = np.float32(0.5)
    exp = oracle.median_filter(oracle.GEOM_GPU, src, L, dr, cb, dst_init=np.full((T, F), SENT, np.float32))
    o = torch.full((T, F), float(SENT), dtype=torch.float32, device="cuda")
    zen.MedianFilterGPU(T, F, L, dr, bool(cb)).filter(torch.from_numpy(src).cuda(), o)
    assert np.array_equal(o.cpu().numpy().view(np.uint32), exp.view(np.uint32))


def test_median_on_reference_s_mag(torch, zen):
    """bit-exact on identical magnitude inputs: the reference's own s_mag -> its harmonic / percussive matrices"""
    n = 0
    for f in sorted(glob.glob(os.path.join(GOLD, "ref_gpu_stages_*.npz"))):
        d = np.load(f)
        be, fs, hop, beta, flags, caus, cb, sse, soft, n_hops, seed = d["params"]
        if sse:
            continue
        nwin, nfft, l_harm, l_perc, lag, W, cola = d["geom"]
        s_mag = torch.from_numpy(d["final_s_mag"]).cuda()
        for name, L, dr in (("final_harmonic_matrix", int(l_harm), int(caus)), ("final_percussive_matrix", int(l_perc), 2)):
            o = torch.zeros_like(s_mag)
            zen.MedianFilterGPU(int(W), int(nfft), L, dr, bool(cb)).filter(s_mag, o)
            assert np.array_equal(o.cpu().numpy().view(np.uint32), d[name].view(np.uint32)), (f, name)
            n += 1
    assert n >= 8


# --------------------------------------------------------------------- box ---

def test_box_vs_npp(torch, zen):
    d = np.load(os.path.join(GOLD, "npp_box.npz"))
    keys = sorted(set(re.match(r"(T\d+_F\d+_L\d+_d\d_[a-z]+)", k).group(1) for k in d.files))
    n = 0
    for k in keys:
        T, F, L, dr = map(int, re.match(r"T(\d+)_F(\d+)_L(\d+)_d(\d)", k).groups())
        if k + "_zgexception" in d.files:
            with pytest.raises(zen.ZgException):
                zen.BoxFilterGPU(T, F, L, dr)
            continue
        o = torch.full((T, F), float(SENT), dtype=torch.float32, device="cuda")
        zen.BoxFilterGPU(T, F, L, dr).filter(torch.from_numpy(d[k + "_src"]).cuda(), o)
        got, ref = o.cpu().numpy(), d[k + "_dst"]
        assert np.array_equal(np.isfinite(got), np.isfinite(ref)), k
        assert np.array_equal(np.isinf(got), np.isinf(ref)), k
        fin = np.isfinite(ref)
        rel = np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])
        assert rel.max() <= 2e-6, (k, rel.max())
        n += 1
    assert n >= 50


@pytest.mark.parametrize("T,F,L,dr", [(12, 4100, 46, 2), (6, 2049, 7, 2), (40, 300, 11, 0), (40, 300, 12, 1), (9, 5000, 3, 2), (30, 64, 23, 0),
                                      (5, 9000, 187, 2)])
def test_box_vs_oracle_with_inf(torch, zen, oracle, T, F, L, dr):
    """BoxFilterGPU (wrap-padded moving average, box.h:194-213) against the oracle's NPP-geometry restatement, on
    reciprocal-of-power data with zeros in it, i.e. +inf inputs (hps.cu:591-592): an inf inside the window gives inf,
    one that has left it leaves no NaN behind (the kernels never subtract)"""
    rng = np.random.default_rng(T * F + L)
    mag = np.abs(rng.standard_normal((T, F))).astype(np.float32)
    mag[rng.random((T, F)) < 0.02] = 0.0
    with np.errstate(divide="ignore"):
        src = (np.float32(1.0) / (mag * mag)).astype(np.float32)
    ref = oracle.box_filter(oracle.GEOM_GPU, src, L, dr)
    dst = torch.zeros((T, F), dtype=torch.float32, device="cuda")
    zen.BoxFilterGPU(T, F, L, dr).filter(torch.from_numpy(src).cuda(), dst)
    got = dst.cpu().numpy()
    assert not np.isnan(got).any()
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.all(np.abs(got[fin] - ref[fin]) <= 4e-6 * np.abs(ref[fin]) + 1e-30)


def test_filters_on_tall_matrices(torch, zen, oracle):
    """more than 65535 rows (rows used to be grid.y)"""
    T, F = 70001, 24
    rng = np.random.default_rng(3)
    src = np.abs(rng.standard_normal((T, F))).astype(np.float32)
    d = torch.from_numpy(src).cuda()
    for L, dr in ((3, 2), (5, 0)):
        dst = torch.zeros((T, F), dtype=torch.float32, device="cuda")
        zen.MedianFilterGPU(T, F, L, dr, True).filter(d, dst)
        assert np.array_equal(dst.cpu().numpy(), oracle.median_filter(oracle.GEOM_GPU, src, L, dr, True))
        dst.zero_()
        zen.BoxFilterGPU(T, F, L, dr).filter(d, dst)
        ref = oracle.box_filter(oracle.GEOM_GPU, src, L, dr)
        assert np.all(np.abs(dst.cpu().numpy() - ref) <= 4e-6 * np.abs(ref))


# --------------------------------------------------------------------- fft ---

@pytest.mark.parametrize("n", [64, 1024, 4096])
def test_fft_vs_cufft_golden(torch, zen, n):
    """libzen/fftw.test.cu tolerance: 2e-4 absolute on uniform(-1,1) inputs, unnormalised both ways"""
    d = np.load(os.path.join(GOLD, "cufft.npz"))
    x = d["n%d_x" % n]
    for key, back in (("fwd", False), ("inv", True)):
        f = zen.FFTC2CWrapperGPU(n)
        f.fft_vec.copy_(torch.from_numpy(x).cuda())
        f.backward() if back else f.forward()
        got = f.fft_vec.cpu().numpy()
        assert np.abs(got - d["n%d_%s" % (n, key)]).max() <= 2e-4


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_fft_vs_float64(torch, zen, n):
    rng = np.random.default_rng(n)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    for back in (False, True):
        f = zen.FFTC2CWrapperGPU(n)
        f.fft_vec.copy_(torch.from_numpy(x).cuda())
        f.backward() if back else f.forward()
        got = f.fft_vec.cpu().numpy().astype(np.complex128)
        exp = np.fft.ifft(x.astype(np.complex128)) * n if back else np.fft.fft(x.astype(np.complex128))
        assert np.abs(got - exp).max() <= 3e-7 * np.sqrt(n) * np.log2(max(n, 2)) * np.abs(exp).max() / np.sqrt(n) + 1e-6


# --------------------------------------------------------------------- HPR ---

HPR_CASES = [
    # fs, hop, beta, flags, causal, copy_bord, sse, soft, n_hops
    (44100.0, 1024, 2.5, 7, True, True, False, False, 40),
    (44100.0, 1024, 2.5, 2, True, True, False, False, 40),
    (44100.0, 1024, 2.5, 7, True, False, False, False, 40),
    (48000.0, 256, 2.0, 7, True, True, False, False, 100),
    (48000.0, 256, 2.0, 7, True, False, False, False, 100),
    (48000.0, 256, 2.0, 7, False, True, False, False, 100),
    (48000.0, 256, 2.0, 7, False, False, False, False, 100),
    (44100.0, 512, 2.5, 7, True, True, True, True, 60),
    (44100.0, 512, 2.5, 7, True, True, False, True, 60),
    (44100.0, 512, 2.5, 5, True, False, False, False, 60),
    (44100.0, 2048, 2.5, 7, True, True, False, False, 12),
    (44100.0, 2048, 2.5, 7, True, False, False, False, 12),
    (44100.0, 4096, 2.5, 7, True, True, False, False, 10),
    (44100.0, 4096, 2.5, 7, False, False, False, False, 10),
    (48000.0, 128, 2.0, 7, True, True, False, False, 120),
    (48000.0, 64, 2.0, 3, True, True, False, False, 150),
    (48000.0, 32, 2.0, 7, True, True, False, False, 260),
    (48000.0, 32, 2.0, 7, False, False, False, False, 260),
]


def _oracle_run(oracle, fs, hop, beta, flags, causal, cb, sse, soft, audio):
    o = oracle.OracleHPR(oracle.GEOM_GPU, fs, hop, beta, flags, oracle.CAUSAL if causal else oracle.ANTICAUSAL, cb)
    if sse:
        o.use_sse_filter()
    if soft:
        o.use_soft_mask()
    return o, o.run(audio)


def _assert_audio(got, ref, what):
    for name, a, b in zip("HPR", got, ref):
        err, snr = peak_norm_err(a, b)
        assert err <= TOL_ABS and snr >= TOL_SNR, (what, name, err, snr)


@pytest.mark.parametrize("fs,hop,beta,flags,causal,cb,sse,soft,n_hops", HPR_CASES)
def test_hpr_streaming_vs_oracle(torch, zen, oracle, fs, hop, beta, flags, causal, cb, sse, soft, n_hops):
    """HPR<GPU>::process_next_hop, hop by hop, against the oracle's GPU-geometry restatement.
    Hard-mask threshold flips (tests/util.py:flip_aware_compare) are identified bin by bin, must be
    borderline decisions, and must be rare; every other hop has to meet the north-star tolerance."""
    audio = synth_audio(n_hops * hop, seed=hop + 3 * int(cb) + int(causal), fs=int(fs))
    o = oracle.OracleHPR(oracle.GEOM_GPU, fs, hop, beta, flags, oracle.CAUSAL if causal else oracle.ANTICAUSAL, cb)
    h = zen.HPR(fs, hop, beta, flags, 0 if causal else 1, cb)
    assert (h.nwin, h.nfft, h.l_harm, h.l_perc, h.lag, h.stft_width) == (o.nwin, o.nfft, o.l_harm, o.l_perc, o.lag, o.stft_width)
    assert np.float32(h.COLA_factor) == np.float32(o.cola)
    if sse:
        o.use_sse_filter()
        h.use_sse_filter()
    if soft:
        o.use_soft_mask()
        h.use_soft_mask()
    r = flip_aware_compare(h, o, audio, hop, flags, hard_mask=not (sse or soft))
    assert r["margin_ok"], ("mask mismatch on a non-borderline bin", r["worst_margin"])
    assert np.count_nonzero(r["flips"]) <= max(1, n_hops // 10), r["flips"]
    assert r["hops_checked"] >= n_hops * 0.8
    for name, e, snr in zip("HPR", r["err"], r["snr"]):
        assert e <= TOL_ABS and snr >= TOL_SNR, (name, e, snr)
    # the *_out members hold [emitted hop | overlap-add tail] like the reference's (hps.h:195-197)
    if not np.count_nonzero(r["flips"][-2:]):
        for name, mine in (("harmonic_out", h.harmonic_out), ("percussive_out", h.percussive_out), ("residual_out", h.residual_out)):
            ref = o.get(name)
            assert np.abs(mine - ref).max() <= TOL_ABS * max(np.abs(r["ref"]["HPR".index(name[0].upper())]).max(), 1e-30), name
    h.close()


@pytest.mark.parametrize("fs,hop,beta,flags,causal,cb,sse,soft,n_hops", HPR_CASES)
def test_hpr_batch_equals_streaming(torch, zen, fs, hop, beta, flags, causal, cb, sse, soft, n_hops):
    """tiles with halo recompute == hop-by-hop streaming, bit for bit; several streams, ragged tile count"""
    n_hops = n_hops * 3 + 1
    n_streams = 3
    audio = np.stack([synth_audio(n_hops * hop, seed=100 + s, fs=int(fs)) for s in range(n_streams)])
    b = zen.HPRBatch(fs, hop, beta, flags, causal=causal, nocopybord=not cb, sse=sse, soft=soft)
    x = torch.from_numpy(audio).cuda()
    outs = b.process(x)
    torch.cuda.synchronize()
    for s in range(n_streams):
        h = zen.HPR(fs, hop, beta, flags, 0 if causal else 1, cb)
        if sse:
            h.use_sse_filter()
        if soft:
            h.use_soft_mask()
        ref = h.run(audio[s])
        h.close()
        for o in range(3):
            if outs[o] is None or (o == 2 and (sse or soft)):
                continue
            assert np.array_equal(outs[o][s].cpu().numpy(), ref[o]), (s, o)
    b.close()


def test_hpr_vs_reference_gpu_golden(torch, zen):
    """separated audio vs the UNMODIFIED reference GPU path (deterministic configs, stft_width == 2)"""
    files = sorted(glob.glob(os.path.join(GOLD, "ref_gpu_audio_*.npz")))
    assert len(files) >= 4
    for f in files:
        d = np.load(f)
        assert d["deterministic"].all()
        be, fs, hop, beta, flags, caus, cb, sse, soft, n_hops, seed = d["params"]
        hop, n_hops = int(hop), int(n_hops)
        audio = synth_audio(n_hops * hop, seed=int(seed), fs=int(fs))
        assert hashlib.sha256(audio.tobytes()).digest() == d["audio_sha"].tobytes()
        h = zen.HPR(float(fs), hop, float(beta), int(flags), int(caus), bool(cb))
        got = h.run(audio)
        h.close()
        _assert_audio(got, [d["harmonic"], d["percussive"], d["residual"]], os.path.basename(f))


def test_hpr_vs_reference_norace_audio(torch, zen):
    """END-TO-END audio of the reference GPU path (thrust + cuFFT + NPP) at the headline hops 256 / 512 / 1024, from the
    reference built with the one-hunk fix of its racy ring shift (oracle/Makefile ref_norace, oracle/ref/norace.patch;
    goldens by oracle/ref/probe_ref_norace.py on a B200, two runs bit-compared).  Hard-mask flips are identified with the
    reference's own per-hop masks; every other hop must meet 1e-4 / 80 dB."""
    files = sorted(glob.glob(os.path.join(GOLD, "ref_norace_*.npz")))
    assert len(files) >= 3, "tests/golden/ref_norace_*.npz missing"
    for f in files:
        d = np.load(f)
        fs, hop, beta, flags, cb, n_hops, seed = d["params"][:7]
        sse, soft = (int(d["params"][7]), int(d["params"][8])) if len(d["params"]) > 7 else (0, 0)
        hop, n_hops, flags = int(hop), int(n_hops), int(flags)
        audio = synth_audio(n_hops * hop, seed=int(seed), fs=int(fs))
        assert hashlib.sha256(audio.tobytes()).digest() == d["audio_sha"].tobytes()
        h = zen.HPR(float(fs), hop, float(beta), flags, 0, bool(cb))
        if sse:
            h.use_sse_filter()
        if soft:
            h.use_soft_mask()
        row = h.stft_width - h.lag
        a_dev = torch.from_numpy(audio).cuda()
        tmp = [torch.zeros(hop, dtype=torch.float32, device="cuda") for _ in range(3)]
        got = [np.zeros(n_hops * hop, np.float32) for _ in range(3)]
        flips = np.zeros(n_hops, dtype=np.int64)
        for i in range(n_hops):
            h.process_hop_io(a_dev[i * hop:].data_ptr(), tmp[0].data_ptr(), tmp[1].data_ptr(), tmp[2].data_ptr())
            h.synchronize()
            for o in range(3):
                got[o][i * hop:(i + 1) * hop] = tmp[o].cpu().numpy()
            if not (sse or soft):      # hard masks: a borderline bin may land on the other side of the threshold
                m = h.materialize()
                for nm in ("harmonic_mask", "percussive_mask"):
                    mine = np.packbits(m[nm][row] != 0)
                    flips[i] += int(np.unpackbits(mine ^ d[nm + "_bits"][i]).sum())
        h.close()
        clean = np.ones(n_hops, dtype=bool)
        for i in np.nonzero(flips)[0]:
            clean[i:i + 2] = False
        assert np.count_nonzero(flips) <= max(2, n_hops // 20), (os.path.basename(f), flips)
        for o, nm in enumerate(("harmonic", "percussive", "residual")):
            ref = d[nm]
            pk = max(float(np.abs(ref).max()), 1e-30)
            g = got[o].reshape(n_hops, hop)[clean]
            r = ref.reshape(n_hops, hop)[clean]
            err = float(np.abs(g - r).max() / pk)
            den = float(np.sum((g.astype(np.float64) - r) ** 2))
            snr = float("inf") if den == 0 else 10 * np.log10(float(np.sum(r.astype(np.float64) ** 2)) / den)
            assert err <= TOL_ABS and snr >= TOL_SNR, (os.path.basename(f), nm, err, snr, int(flips.sum()))


def test_hpr_materialize_vs_oracle(torch, zen, oracle):
    """the reference's public stft_width x nfft matrices, rebuilt on demand"""
    for (fs, hop, beta, flags, causal, cb, sse, soft, n_hops) in [HPR_CASES[0], HPR_CASES[2], HPR_CASES[6], HPR_CASES[7]]:
        audio = synth_audio(n_hops * hop, seed=5, fs=int(fs))
        o, _ = _oracle_run(oracle, fs, hop, beta, flags, causal, cb, sse, soft, audio)
        h = zen.HPR(fs, hop, beta, flags, 0 if causal else 1, cb)
        if sse:
            h.use_sse_filter()
        if soft:
            h.use_soft_mask()
        h.run(audio)
        m = h.materialize()
        h.close()
        stft_o = o.get("sliding_stft")
        scale = np.abs(stft_o).max()
        assert np.abs(m["sliding_stft"] - stft_o).max() <= 2e-6 * scale
        smag_o = o.get("s_mag")
        assert np.abs(m["s_mag"] - smag_o).max() <= 4e-6 * smag_o.max()
        if not sse:
            # medians are selections: identical up to the tiny FFT differences of the selected element
            for nm in ("harmonic_matrix", "percussive_matrix"):
                assert np.abs(m[nm] - o.get(nm)).max() <= 4e-6 * smag_o.max(), nm
            for nm in ("harmonic_mask", "percussive_mask", "residual_mask"):
                assert np.mean(m[nm] != o.get(nm)) <= 2e-3, nm


def test_hps_test_cu_properties(torch, zen):
    """libzen/hps.test.cu:160-372 restated: outputs != input, copybord != nocopybord on the GPU,
    percussive-only leaves harmonic/residual at 0, reset_buffers() reproduces the first hop bit for bit"""
    fs, hop, n_hops = 48000.0, 256, 100
    rng = np.random.default_rng(0)
    data = rng.uniform(-1, 1, n_hops * hop).astype(np.float32)
    a = zen.HPR(fs, hop, 2.0, 7, 0, True)
    b = zen.HPR(fs, hop, 2.0, 7, 0, False)
    oa, ob = a.run(data), b.run(data)
    for o in oa:
        assert not np.array_equal(o, data)
    assert any(not np.array_equal(x, y) for x, y in zip(oa, ob))
    c = zen.HPR(fs, hop, 2.0, zen.OUTPUT_PERCUSSIVE, 0, True)
    oc = c.run(data)
    assert np.all(oc[0] == 0) and np.all(oc[2] == 0) and np.any(oc[1] != 0)
    assert np.all(c.harmonic_out == 0) and np.all(c.residual_out == 0)
    first = a.run(data[:hop], 1)
    a.reset_buffers()
    x1 = a.run(data[:hop], 1)
    a.reset_buffers()
    x2 = a.run(data[:hop], 1)
    assert all(np.array_equal(p, q) for p, q in zip(x1, x2))
    assert not all(np.array_equal(p, q) for p, q in zip(first, x1))
    for h in (a, b, c):
        h.close()


def test_realtime_api_and_io(torch, zen, oracle):
    """HPRRealtime + IOGPU driven exactly like zen/fakert.h:217-251 (mapped in, mapped out)"""
    fs, hop, n_hops = 44100.0, 1024, 30
    audio = synth_audio(n_hops * hop, seed=9)
    _, ref = _oracle_run(oracle, fs, hop, 2.5, 2, True, True, False, False, audio)
    hp = zen.HPRRealtime(fs, hop, 2.5, zen.OUTPUT_PERCUSSIVE)
    io = zen.IOGPU(hop)
    hp.warmup(io, test_iters=20)
    out = np.zeros_like(audio)
    for i in range(n_hops):
        io.host_in[:] = audio[i * hop:(i + 1) * hop]
        hp.process_next_hop(io.device_in)
        hp.copy_percussive(io.device_out)
        out[i * hop:(i + 1) * hop] = io.host_out
    err, snr = peak_norm_err(out, ref[1])
    assert err <= TOL_ABS and snr >= TOL_SNR, (err, snr)
    io.close()


def test_constructor_errors(torch, zen):
    with pytest.raises(zen.ZgException):
        zen.HPRIOffline(44100.0, 4096, 300, 2.0, 2.0)     # hps.cu:33-36
    with pytest.raises(zen.ZgException):
        zen.MedianFilterGPU(9, 9, 10, 0)                  # mfilt.test.cu:235-244
    with pytest.raises(zen.ZgException):
        zen.MedianFilterGPU(9, 9, 10, 2, True)
    with pytest.raises(zen.ZgException):
        zen.BoxFilterGPU(9, 9, 10, 1)


# ----------------------------------------------------------------- offline ---

@pytest.mark.parametrize("nocb,sse,soft", [(False, False, False), (True, False, False), (False, False, True), (False, True, False)])
def test_offline_vs_oracle(torch, zen, oracle, nocb, sse, soft):
    """HPRIOffline<GPU>::process: padded sizes as in hps_gpu_public.test.cu:60-80 (n not a multiple of the hop)"""
    n = 10 * 4096 + 11
    audio = synth_audio(n, seed=21)
    ref = oracle.offline_process(oracle.GEOM_GPU, 44100.0, 4096, 256, 2.5, 2.5, audio, nocopybord=nocb, sse=sse, soft=soft)
    off = zen.HPRIOffline(44100.0, 4096, 256, 2.5, 2.5, nocopybord=nocb)
    if sse:
        off.use_sse_filter()
    if soft:
        off.use_soft_mask()
    got = off.process(audio)
    assert all(g.size == n for g in got)
    assert np.all(got[2] == 0)                 # reference quirk: residual is all zeros (hps.cu:45-48, 200-204)
    _assert_audio(got[:2], ref[:2], "offline")


def test_offline_vs_reference_gpu_golden(torch, zen):
    """pass 1 (hop 4096, stft_width 2) of the reference GPU path is deterministic: harmonic must match"""
    audio = synth_audio(10 * 4096 + 11, seed=21)
    for name, nocb in (("ref_offline_gpu.npz", False), ("ref_offline_gpu_nocb.npz", True)):
        d = np.load(os.path.join(GOLD, name))
        got = zen.HPRIOffline(44100.0, 4096, 256, 2.5, 2.5, nocopybord=nocb).process(audio)
        err, snr = peak_norm_err(got[0], d["harmonic"])
        assert err <= TOL_ABS and (snr >= TOL_SNR or np.all(d["harmonic"] == 0)), (name, err, snr)
        assert int(d["residual_all_zero"]) == 1 and np.all(got[2] == 0)


def test_batch_host_path(torch, zen):
    """host-buffer entry point (chunked H2D / compute / D2H) == device-resident entry point"""
    fs, hop, n_hops, n_streams = 44100.0, 1024, 50, 5
    audio = np.stack([synth_audio(n_hops * hop, seed=300 + s) for s in range(n_streams)])
    b = zen.HPRBatch(fs, hop, 2.5, zen.OUTPUT_PERCUSSIVE)
    dev = b.process(torch.from_numpy(audio).cuda())[1].cpu().numpy()
    host_out = np.zeros_like(audio)
    b.process_host(audio, [None, host_out, None])
    assert np.array_equal(dev, host_out)
    assert b.last_launches >= 1
    b.close()


@pytest.mark.parametrize("fs,hop,flags,causal,cb", [(44100.0, 1024, 7, True, True), (44100.0, 1024, 7, True, False),
                                                     (48000.0, 256, 7, False, True), (48000.0, 256, 3, False, False),
                                                     (44100.0, 4096, 7, True, True), (44100.0, 2048, 6, True, False),
                                                     (44100.0, 512, 2, True, True)])
def test_decision_path_equals_median_path(torch, zen, fs, hop, flags, causal, cb):
    """hard masks decided by counting taps against the exact threshold == selecting the median and comparing:
    the separated audio must be identical bit for bit"""
    n_hops, n_streams = 60, 2
    audio = np.stack([synth_audio(n_hops * hop, seed=700 + s, fs=int(fs)) for s in range(n_streams)])
    x = torch.from_numpy(audio).cuda()
    res = []
    for no_decide in (False, True):
        if no_decide:
            os.environ["ZEN_B200_NO_DECIDE"] = "1"
        else:
            os.environ.pop("ZEN_B200_NO_DECIDE", None)
        b = zen.HPRBatch(fs, hop, 2.5, flags, causal=causal, nocopybord=not cb)
        outs = b.process(x)
        torch.cuda.synchronize()
        res.append([o.cpu().numpy() if o is not None else None for o in outs])
        b.close()
    os.environ.pop("ZEN_B200_NO_DECIDE", None)
    for a, b_ in zip(*res):
        if a is not None:
            assert np.array_equal(a, b_)
            assert np.abs(a).max() > 0


@pytest.mark.parametrize("fs,hop,flags", [(44100.0, 1024, 2), (44100.0, 1024, 7), (44100.0, 1024, 4), (44100.0, 1024, 5), (44100.0, 512, 2),
                                          (44100.0, 512, 3), (44100.0, 2048, 7), (44100.0, 4096, 6), (48000.0, 1024, 2), (22050.0, 1024, 7),
                                          (44100.0, 128, 7)])
def test_fast_tile_kernel_equals_general_kernel(torch, zen, fs, hop, flags):
    """hpr_tile_fast_kernel (window and overlap-add fused into the FFT stages, eight bins decided per thread) against
    hpr_tile_kernel (ZEN_B200_NO_FAST=1) on the plans it serves: identical bit for bit, ragged tiles, odd stream count"""
    n_hops, n_streams = 131, 3
    audio = np.stack([synth_audio(n_hops * hop, seed=900 + s, fs=int(fs)) for s in range(n_streams)])
    audio[1, : 7 * hop] = 0.0     # a silent lead-in: zero magnitudes, trivial thresholds
    x = torch.from_numpy(audio).cuda()
    res = []
    for no_fast in (False, True):
        if no_fast:
            os.environ["ZEN_B200_NO_FAST"] = "1"
        else:
            os.environ.pop("ZEN_B200_NO_FAST", None)
        b = zen.HPRBatch(fs, hop, 2.5, flags)
        outs = b.process(x)
        torch.cuda.synchronize()
        res.append([o.cpu().numpy() if o is not None else None for o in outs])
        b.close()
    os.environ.pop("ZEN_B200_NO_FAST", None)
    for a, b_ in zip(*res):
        if a is not None:
            assert np.array_equal(a, b_)
            assert np.isfinite(a).all() and np.abs(a).max() > 0


@pytest.mark.parametrize("fs,hop,flags,cb,sse,soft", [(44100.0, 1024, 2, True, False, False), (44100.0, 1024, 7, False, False, False),
                                                      (48000.0, 256, 7, True, False, False), (44100.0, 512, 3, True, True, False),
                                                      (44100.0, 512, 7, True, False, True), (44100.0, 4096, 7, True, False, False)])
def test_resident_realtime_kernel_equals_per_hop_launches(torch, zen, fs, hop, flags, cb, sse, soft):
    """zen_hpr_realtime_begin: the persistent kernel (state in shared memory, doorbell in mapped memory) must
    reproduce the per-launch path bit for bit, survive an idle time-out, and hand its state back on pause"""
    import time
    n_hops = 40
    audio = synth_audio(n_hops * hop, seed=77, fs=int(fs))

    def make():
        h = zen.HPR(fs, hop, 2.5, flags, 0, cb)
        if sse:
            h.use_sse_filter()
        if soft:
            h.use_soft_mask()
        return h
    ref_obj = make()
    ref = ref_obj.run(audio)
    h = make()
    io = zen.IOGPU(hop)
    outs = [zen.IOGPU(hop) for _ in range(3)]
    got = [np.zeros(n_hops * hop, np.float32) for _ in range(3)]
    h.realtime_begin()
    for i in range(n_hops):
        if i == 10:
            time.sleep(0.6)                  # longer than the idle time-out: the kernel leaves and is brought back
        if i == 20:
            mid_state = h.percussive_out     # pauses the resident kernel, reads the state from global memory
            assert np.array_equal(mid_state[:hop], got[1][(i - 1) * hop:i * hop])
        if i == 30:
            h.realtime_end()                 # continue with per-hop launches on the same object
        io.host_in[:] = audio[i * hop:(i + 1) * hop]
        h.process_hop_io(io.device_in, outs[0].device_out, outs[1].device_out, outs[2].device_out)
        h.synchronize()
        for o in range(3):
            if flags & (1 << o):
                got[o][i * hop:(i + 1) * hop] = outs[o].host_out
    for o in range(3):
        if flags & (1 << o):
            assert np.array_equal(got[o], ref[o]), o
    assert np.array_equal(h.percussive_out, ref_obj.percussive_out)
    h.close()
    ref_obj.close()

@pytest.mark.gpu
@pytest.mark.parametrize("hop,flags,cb,soft", [(1024, 2, True, False), (512, 7, True, False), (1024, 7, False, False), (512, 7, True, True)])
def test_resident_kernel_two_call_sequence(torch, zen, hop, flags, cb, soft):
    """HPRRealtime's own sequence (process_next_hop, then copy_*; zen/fakert.h:229-230) on a resident session: the hop
    is only submitted, copy_* waits for the tagged output and unpacks it on the host (host-visible destination) or lets
    the kernel copy it (device memory).  Bit-identical to the per-launch path, also across an idle time-out and when a
    copy is skipped or repeated."""
    import time
    fs, n_hops = 44100.0, 30
    audio = synth_audio(n_hops * hop, seed=79)

    def make():
        h = zen.HPR(fs, hop, 2.5, flags, 0, cb)
        if soft:
            h.use_soft_mask()
        return h
    ref_obj = make()
    ref = ref_obj.run(audio)
    ref_obj.close()
    h = make()
    h.realtime_begin()
    io = zen.IOGPU(hop)
    houts = [zen.IOGPU(hop) for _ in range(3)]
    douts = [torch.zeros(hop, dtype=torch.float32, device="cuda") for _ in range(3)]
    L = zen._lib.lib() if hasattr(zen, "_lib") else None
    from zen_b200 import _lib
    L = _lib.lib()
    copies = [L.zen_hpr_copy_harmonic, L.zen_hpr_copy_percussive, L.zen_hpr_copy_residual]
    got = [np.zeros(n_hops * hop, np.float32) for _ in range(3)]
    for i in range(n_hops):
        sl = slice(i * hop, (i + 1) * hop)
        if i == 12:
            time.sleep(0.6)                      # idle time-out between two hops
        io.host_in[:] = audio[sl]
        h.process_next_hop(io.device_in)
        if i == 17:
            time.sleep(0.6)                      # idle time-out between the submission and the copy
        for o in range(3):
            if not flags & (1 << o):
                continue
            if i % 5 == 3:                       # device-memory destination: the kernel copies
                assert copies[o](h._h, douts[o].data_ptr()) == 0
                torch.cuda.synchronize()
                got[o][sl] = douts[o].cpu().numpy()
            else:
                assert copies[o](h._h, houts[o].device_out) == 0
                if i % 7 == 2:                   # asking twice gives the same hop
                    assert copies[o](h._h, houts[o].device_out) == 0
                got[o][sl] = houts[o].host_out
    for o in range(3):
        if flags & (1 << o):
            assert np.array_equal(got[o], ref[o]), o
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("hop,flags,mode", [(1024, 7, "device"), (1024, 2, "pull"), (32, 7, "push"), (2048, 3, "device"),
                                            (4096, 2, "mixed")])
def test_resident_kernel_request_protocols(torch, zen, hop, flags, mode):
    """the resident kernel takes the hop either pushed in tagged 16-byte groups (host-visible buffers, the IOGPU
    case) or pulled by the kernel itself (device memory, or ZEN_B200_RT_PUSH=0), and emits either tagged groups or
    plain stores + completion flag: every combination must give the per-launch result bit for bit, also when the
    buffers change from one hop to the next"""
    fs, n_hops = 44100.0, 24
    audio = synth_audio(n_hops * hop, seed=78)
    ref_obj = zen.HPR(fs, hop, 2.5, flags, 0, True)
    ref = ref_obj.run(audio)
    ref_obj.close()
    if mode == "pull":
        os.environ["ZEN_B200_RT_PUSH"] = "0"
    try:
        h = zen.HPR(fs, hop, 2.5, flags, 0, True)
        h.realtime_begin()
        ios = [zen.IOGPU(hop) for _ in range(2)]
        houts = [[zen.IOGPU(hop) for _ in range(3)] for _ in range(2)]
        d_in = torch.from_numpy(audio).cuda()
        d_out = [torch.zeros(n_hops * hop, dtype=torch.float32, device="cuda") for _ in range(3)]
        got = [np.zeros(n_hops * hop, np.float32) for _ in range(3)]
        for i in range(n_hops):
            sl = slice(i * hop, (i + 1) * hop)
            dev_in = mode == "device" or (mode == "mixed" and i % 3 == 0)
            dev_out = mode == "device" or (mode == "mixed" and i % 4 < 2)
            io, ho = ios[i & 1], houts[(i >> 1) & 1]
            io.host_in[:] = audio[sl]
            src = d_in[sl].data_ptr() if dev_in else io.device_in
            dst = [d_out[o][sl].data_ptr() if dev_out else ho[o].device_out for o in range(3)]
            h.process_hop_io(src, *dst)
            if dev_out:
                torch.cuda.synchronize()
            for o in range(3):
                if flags & (1 << o):
                    got[o][sl] = d_out[o][sl].cpu().numpy() if dev_out else ho[o].host_out
        for o in range(3):
            if flags & (1 << o):
                assert np.array_equal(got[o], ref[o]), o
        assert max(np.abs(got[o]).max() for o in range(3)) > 0
        h.close()
    finally:
        os.environ.pop("ZEN_B200_RT_PUSH", None)



def test_batch_many_streams_sampled(torch, zen):
    """a batch large enough to keep every resident CTA busy for several work items (work queue, L2-resident scratch
    reuse between items): sampled streams must equal the per-hop path bit for bit"""
    fs, hop, n_hops, n_streams = 44100.0, 1024, 150, 700
    base = np.stack([synth_audio(n_hops * hop, seed=900 + s) for s in range(8)])
    audio = np.tile(base, (n_streams // 8 + 1, 1))[:n_streams].copy()
    audio *= (1.0 - 0.3 * (np.arange(n_streams) % 7)[:, None] / 7.0).astype(np.float32)
    b = zen.HPRBatch(fs, hop, 2.5, 7)
    outs = b.process(torch.from_numpy(audio).cuda())
    torch.cuda.synchronize()
    for s in (0, 1, 349, 698, 699):
        h = zen.HPR(fs, hop, 2.5, 7, 0, True)
        ref = h.run(audio[s])
        h.close()
        for o in range(3):
            assert np.array_equal(outs[o][s].cpu().numpy(), ref[o]), (s, o)
    b.close()


SWEEP = [
    # fs, hop, beta, flags, causal, copy_bord, sse, soft, n_hops
    (22050.0, 512, 2.0, 7, True, True, False, False, 50),
    (96000.0, 1024, 3.0, 7, True, True, False, False, 40),
    (44100.0, 1024, 1.0, 7, True, True, False, False, 40),     # beta - eps < 1 <= beta
    (44100.0, 512, 0.5, 7, True, True, False, False, 50),      # beta < 1: both masks can be 1, residual mask -1
    (44100.0, 1024, 2.5, 1, True, True, False, False, 40),     # harmonic only
    (44100.0, 1024, 2.5, 4, True, True, False, False, 40),     # residual only: masks of disabled outputs are 0 -> Mr = 1
    (44100.0, 1024, 2.5, 5, True, False, False, False, 40),    # H + R, no copy border
    (44100.0, 256, 2.5, 6, False, True, False, False, 80),     # P + R, anticausal
    (44100.0, 1024, 0.7, 3, True, True, False, True, 40),      # soft mask with (int)beta == 0: x^0 / (x^0 + y^0 + eps)
    (44100.0, 2048, 2.5, 3, True, True, False, True, 12),      # soft mask, L = 93 (warp-resident sliding window)
    (44100.0, 4096, 2.5, 2, False, True, False, True, 8),      # soft mask, L = 187
    (44100.0, 2048, 2.5, 7, False, True, True, False, 12),     # SSE at a large hop
    (8000.0, 256, 2.0, 7, True, True, False, False, 60),
]


@pytest.mark.parametrize("fs,hop,beta,flags,causal,cb,sse,soft,n_hops", SWEEP)
def test_hpr_parameter_sweep_vs_oracle(torch, zen, oracle, fs, hop, beta, flags, causal, cb, sse, soft, n_hops):
    """sample rates, beta below / at / above 1, every output-flag subset, long frequency windows"""
    audio = synth_audio(n_hops * hop, seed=int(fs) % 97 + hop, fs=int(fs))
    try:
        o = oracle.OracleHPR(oracle.GEOM_GPU, fs, hop, beta, flags, oracle.CAUSAL if causal else oracle.ANTICAUSAL, cb)
    except ValueError:
        with pytest.raises(zen.ZgException):
            zen.HPR(fs, hop, beta, flags, 0 if causal else 1, cb)
        return
    h = zen.HPR(fs, hop, beta, flags, 0 if causal else 1, cb)
    assert (h.l_harm, h.l_perc, h.lag, h.stft_width) == (o.l_harm, o.l_perc, o.lag, o.stft_width)
    if sse:
        o.use_sse_filter()
        h.use_sse_filter()
    if soft:
        o.use_soft_mask()
        h.use_soft_mask()
    r = flip_aware_compare(h, o, audio, hop, flags, hard_mask=not (sse or soft))
    assert r["margin_ok"], r["worst_margin"]
    assert np.count_nonzero(r["flips"]) <= max(2, n_hops // 8), r["flips"]
    for name, e, snr in zip("HPR", r["err"], r["snr"]):
        assert e <= TOL_ABS and snr >= TOL_SNR, (name, e, snr)
    h.close()


@pytest.mark.parametrize("kind", ["zeros", "dc", "impulse", "alternating", "tiny", "zeros_then_noise"])
@pytest.mark.parametrize("sse,soft", [(False, False), (False, True), (True, False)])
def test_hpr_edge_inputs(torch, zen, oracle, kind, sse, soft):
    """silence (0/0 in the masks, 1/0 in the SSE reciprocal), DC, a single impulse, Nyquist, near-denormal levels"""
    fs, hop, n_hops = 44100.0, 512, 40
    n = n_hops * hop
    rng = np.random.default_rng(3)
    x = {"zeros": np.zeros(n), "dc": np.full(n, 0.5), "impulse": np.eye(1, n, 7 * hop + 13)[0],
         "alternating": np.where(np.arange(n) % 2 == 0, 1.0, -1.0), "tiny": 1e-30 * rng.standard_normal(n),
         "zeros_then_noise": np.concatenate([np.zeros(n // 2), 0.3 * rng.standard_normal(n - n // 2)])}[kind].astype(np.float32)
    o = oracle.OracleHPR(oracle.GEOM_GPU, fs, hop, 2.5, 7, oracle.CAUSAL, True)
    h = zen.HPR(fs, hop, 2.5, 7, 0, True)
    if sse:
        o.use_sse_filter()
        h.use_sse_filter()
    if soft:
        o.use_soft_mask()
        h.use_soft_mask()
    with np.errstate(all="ignore"):
        r = flip_aware_compare(h, o, x, hop, 7, hard_mask=not (sse or soft))
    for name, g, ref in zip("HPR", r["got"], r["ref"]):
        assert np.array_equal(np.isnan(g), np.isnan(ref)), (name, "NaN pattern")
    if kind == "zeros":
        for g in r["got"]:
            assert np.all(np.nan_to_num(g) == 0)
    assert r["margin_ok"], r["worst_margin"]
    for name, e, snr, ref in zip("HPR", r["err"], r["snr"], r["ref"]):
        if np.isfinite(e) and np.nanmax(np.abs(ref)) > 0:
            assert e <= TOL_ABS and (snr >= TOL_SNR or not np.isfinite(snr)), (name, e, snr)
    h.close()
