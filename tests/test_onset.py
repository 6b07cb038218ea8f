"""The consumer of the percussive output: BTrack's onset detection function
(demos/beat-tracking/OnsetDetection.cpp) on device buffers.  The oracle restates it literally - including the in-place
windowing that leaves the older half of the frame zero - and is pinned by the reference's own source compiled unmodified
(oracle/Makefile ref_odf -> tests/golden/odf_csd.npz, generator oracle/ref/make_odf_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oraclebind as oracle
from tests.odf_inputs import CASES, make_input

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "odf_csd.npz")
# float arithmetic through atan2f / cosf / a 512-bin sum: the oracle differs from the reference build by its window
# (libm cosf against gcem's compile-time cosine, an ulp on 220 of 512 entries) - 3e-6 of the largest sample observed
REL_TOL = 2e-5


def close(a, b, scale=None):
    """The reference takes sqrtf(m^2 + p^2 - 2 m p cos(dev)) (OnsetDetection.cpp:115-117); when a bin barely changes,
    rounding can make the argument slightly negative and the whole ODF sample NaN - in the reference too (bug as spec).
    Whether that happens on a given hop hangs on the last bit of cosf, so hops where either side is NaN are set aside
    (they must be rare) and the others compared."""
    ok = np.isfinite(a) & np.isfinite(b)
    assert ok.mean() >= 0.98, ok.mean()
    if not ok.any():
        return True
    scale = max(float(np.abs(b[ok]).max()), 1e-30) if scale is None else scale
    return float(np.abs(a[ok] - b[ok]).max()) <= REL_TOL * scale


def test_onset_oracle_pinned_by_the_reference_source():
    g = np.load(GOLD)
    w = oracle.onset_window()
    assert np.abs(w - g["window"]).max() <= 1.2e-7
    for name, n_hops, kind, arg in CASES:
        y = oracle.onset_csd(make_input(n_hops, kind, arg))
        ref = g[name]
        assert y.shape == ref.shape
        assert np.abs(y - ref).max() <= REL_TOL * max(float(ref.max()), 1e-30), name


def test_onset_frame_is_the_current_hop_only():
    """the reference windows its frame in place, so the half it shifts down is always the zero half it started with:
    the ODF sample of a hop depends on the last three hops only (what lets the device kernel start anywhere)"""
    x = make_input(40, "synth", 5)
    full = oracle.onset_csd(x)
    # the same hops after a different past give the same samples from the third hop on
    y = x.copy()
    y[:20 * 256] = make_input(20, "noise", 1)
    other = oracle.onset_csd(y)
    assert np.array_equal(full[22:], other[22:])
    assert not np.array_equal(full[20:22], other[20:22])


@pytest.mark.gpu
def test_onset_kernel_vs_oracle_and_reference():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zen_b200 import hps
    g = np.load(GOLD)
    odf = hps.OnsetDetectionFunction()
    for name, n_hops, kind, arg in CASES:
        x = make_input(n_hops, kind, arg)
        got = odf.calculate_samples(torch.from_numpy(x).cuda()).cpu().numpy()
        ref, orc = g[name], oracle.onset_csd(x)
        scale = max(float(ref.max()), 1e-30)
        assert got.shape == ref.shape
        assert np.abs(got - orc).max() <= REL_TOL * scale, (name, np.abs(got - orc).max(), scale)
        assert np.abs(got - ref).max() <= REL_TOL * scale, name


@pytest.mark.gpu
def test_onset_kernel_batched_streams_and_tiles():
    """many streams per launch, long enough that a stream is cut into tiles (each re-analyses two hops), rows in a wider
    buffer; fed straight from the percussive output of the batched HPR kernel at hop 256"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zen_b200 import hps
    from zen_b200.synth import synth_audio
    n_streams, n_hops = 5, 3000
    x = np.stack([synth_audio(n_hops * 256 + 100, seed=40 + s) for s in range(n_streams)])
    d = torch.from_numpy(x).cuda()
    odf = hps.OnsetDetectionFunction()
    got = odf.calculate_samples(d[:, :n_hops * 256]).cpu().numpy()          # stride 256 n_hops + 100
    assert got.shape == (n_streams, n_hops)
    for s in range(n_streams):
        orc = oracle.onset_csd(x[s, :n_hops * 256])
        assert close(got[s], orc), s
    # HPR percussive output -> ODF without leaving the device
    b = hps.HPRBatch(44100.0, 256, 2.5, hps.OUTPUT_PERCUSSIVE)
    _, p, _ = b.process(d[:, :400 * 256].contiguous())
    o = odf.calculate_samples(p).cpu().numpy()
    pc = p.cpu().numpy()
    for s in range(n_streams):
        orc = oracle.onset_csd(pc[s])
        assert close(o[s], orc), s
    assert odf.calculate_samples(d[:, :0]).shape == (n_streams, 0)
