"""Downstream consumer of the harmonic output: the McLeod pitch method of
demos/pitch-tracking (pitch.cpp:40-135), on device buffers.

* the oracle (oracle/hpr_oracle.c:zo_mpm_pitch) is PINNED: bit-identical to the UNMODIFIED pitch.cpp compiled from
  /root/reference (oracle/Makefile ref_mpm -> tests/golden/mpm_pitch.npz, oracle/ref/make_mpm_golden.py);
* the CUDA kernel (csrc/mpm.cu) against the oracle: autocorrelation within FFT rounding, the same pitch / -1 decisions;
* fed straight from the device buffer the batched HPR kernel wrote its harmonic hops to."""
import os

import numpy as np
import pytest

from tests.mpm_inputs import CASES, make_input

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpm_pitch.npz")


def test_mpm_oracle_pinned_by_the_reference_pitch_cpp(oracle):
    g = np.load(GOLD)
    assert len(g.files) == len(CASES)
    for name, n, fs, kind, arg in CASES:
        got = np.float32(oracle.mpm_pitch(make_input(n, fs, kind, arg), fs))
        ref = g[name]
        assert got == ref or (np.isnan(got) and np.isnan(ref)), (name, got, ref)


def test_mpm_oracle_known_behaviour(oracle):
    """what the reference's MPM does as written: -1 below 80 Hz and on silence; the half-applied power spectrum
    (pitch.cpp:50-52) biases the estimate low - restated, not repaired"""
    fs = 44100.0
    assert oracle.mpm_pitch(np.zeros(4096, np.float32), fs) == -1.0
    assert oracle.mpm_pitch(make_input(4096, fs, "tone", 60.0), fs) == -1.0
    p = oracle.mpm_pitch(make_input(4096, fs, "tone", 220.0), fs)
    assert 150.0 < p < 230.0


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a GPU")
    return t


@pytest.mark.gpu
def test_mpm_kernel_vs_oracle(torch, oracle):
    from zen_b200 import hps
    for n in (256, 1024, 4096):
        cases = [c for c in CASES if c[1] == n and c[2] == 44100.0]
        x = np.stack([make_input(n, fs, kind, arg) for _, _, fs, kind, arg in cases])
        m = hps.MPM(n, 44100.0)
        got, nsdf = m.pitch(torch.from_numpy(x).cuda(), want_nsdf=True)
        got, nsdf = got.cpu().numpy(), nsdf.cpu().numpy()
        for i, (name, _, fs, kind, arg) in enumerate(cases):
            ref, rn = oracle.mpm_pitch(x[i], fs, want_nsdf=True)
            scale = max(float(np.abs(rn).max()), 1e-30)
            assert np.abs(nsdf[i] - rn).max() <= 2e-5 * scale + 1e-9, (name, np.abs(nsdf[i] - rn).max(), scale)
            if ref < 0 or not np.isfinite(ref):
                assert got[i] == ref or (got[i] < 0 and ref < 0), (name, got[i], ref)
            else:
                assert abs(got[i] - ref) <= 2e-4 * ref, (name, got[i], ref)


@pytest.mark.gpu
def test_mpm_fed_from_the_harmonic_output_on_the_device(torch, oracle):
    """demos/pitch-tracking/main.cu:90-107: HPRRealtime(fs, 4096, 2.5, OUTPUT_HARMONIC) hop by hop, MPM on every
    harmonic hop - here the batched kernel writes the harmonic hops to device memory and the pitch kernel reads them
    there: no host round trip between the two"""
    from zen_b200 import hps
    from zen_b200.synth import synth_audio
    fs, hop, n_hops = 44100.0, 4096, 12
    audio = synth_audio(n_hops * hop, seed=5)
    b = hps.HPRBatch(fs, hop, 2.5, hps.OUTPUT_HARMONIC)
    harm = b.process(torch.from_numpy(audio[None]).cuda())[0]          # [1, n_hops * hop] on the device
    pitches = hps.MPM(hop, fs).pitch(harm.reshape(n_hops, hop)).cpu().numpy()
    h_host = harm.cpu().numpy().reshape(n_hops, hop)
    agree = 0
    for i in range(n_hops):
        ref = oracle.mpm_pitch(h_host[i], fs)
        ok = (pitches[i] == ref) or (ref > 0 and np.isfinite(ref) and abs(pitches[i] - ref) <= 2e-4 * ref) or (ref < 0 and pitches[i] < 0)
        agree += bool(ok)
    assert agree == n_hops, (pitches, agree)
    b.close()
