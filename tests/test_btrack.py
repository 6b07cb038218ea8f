"""The beat tracker behind the onset detection function: zen_btrack_* (csrc/btrack.cu, host
code as in the reference) against the reference's BTrack.cpp compiled unmodified (oracle/Makefile ref_btrack ->
tests/golden/btrack.npz, generator oracle/ref/make_btrack_golden.py).  Fed with the onset detection samples the
reference's own tracker consumed, every decision must be the reference's: beat flags, tempo estimates and cumulative
scores bit for bit - including the run whose onset detection function returns a NaN."""
import os

import numpy as np
import pytest

from tests.btrack_inputs import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "btrack.npz")


def test_lookup_tables_match_the_precomputed_ones():
    """BTrackPrecomputed.h ("done with numpy") recomputed from formulas: Rayleigh weighting with parameter 43, Gaussian
    tempo transition matrix with sigma 5"""
    from zen_b200 import hps
    g = np.load(GOLD)
    r, t = hps.BTrack(44100).tables()
    assert np.array_equal(r.view(np.uint32), g["rayleigh"].view(np.uint32))
    assert np.array_equal(t.view(np.uint32), g["transition"].view(np.uint32))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tracker_decisions_are_the_references(case):
    from zen_b200 import hps
    name, fs, n_hops, kind, arg = case
    g = np.load(GOLD)
    odf = g[name + "_odf"]
    assert odf.size == n_hops
    beat, tempo, score = hps.BTrack(int(fs)).process_odf(odf)
    assert np.array_equal(beat, g[name + "_beat"].astype(bool))
    assert np.array_equal(tempo.view(np.uint32), g[name + "_tempo"].view(np.uint32))
    assert np.array_equal(score.view(np.uint32), g[name + "_score"].view(np.uint32))
    assert beat.sum() >= 5


def test_tracker_keeps_its_state_between_calls():
    from zen_b200 import hps
    g = np.load(GOLD)
    odf = g["clicks120_44k_odf"]
    whole = hps.BTrack(44100).process_odf(odf)
    b = hps.BTrack(44100)
    parts = [b.process_odf(odf[i:j]) for i, j in ((0, 1), (1, 700), (700, 701), (701, odf.size))]
    for k in range(3):
        assert np.array_equal(np.concatenate([p[k] for p in parts]).view(np.uint8), whole[k].view(np.uint8))
    assert hps.BTrack(44100).process_odf(odf[:0])[0].size == 0


def test_oracle_onset_samples_match_what_the_reference_tracker_consumed():
    """the onset detection oracle (zo_onset_csd) against BTrack::lastOnset of the same reference build"""
    from oracle import oraclebind as oracle
    from tests.btrack_inputs import make_input
    g = np.load(GOLD)
    for name, fs, n_hops, kind, arg in CASES[:2]:
        y = oracle.onset_csd(make_input(fs, n_hops, kind, arg))
        ref = g[name + "_odf"]
        ok = np.isfinite(y) & np.isfinite(ref)
        assert ok.mean() > 0.98
        # (5e-5 of full scale on one hop of these fixtures: the window's last bit, amplified where a bin barely changes)
        assert np.abs(y[ok] - ref[ok]).max() <= 1e-4 * float(ref[ok].max())


@pytest.mark.gpu
def test_percussive_output_to_beats_on_the_device_and_host():
    """HPR (device) -> onset detection function (device) -> beat tracker (host): shapes and a sane tempo"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zen_b200 import hps
    from tests.btrack_inputs import make_input
    x = make_input(44100.0, 1500, "clicks", 120.0)
    d = torch.from_numpy(x[None]).cuda()
    _, p, _ = hps.HPRBatch(44100.0, 256, 2.5, hps.OUTPUT_PERCUSSIVE).process(d)
    beat, tempo, score = hps.BTrack(44100).process_percussive(p[0])
    assert beat.shape == (1500,) and tempo.shape == (1500,) and score.shape == (1500,)
    assert beat.sum() >= 3 and 60.0 <= float(tempo[-1]) <= 200.0
