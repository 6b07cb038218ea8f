"""Full-length parity of BASELINE.json configs[2], [3] and [4] against the oracle (-m gpu).

VERDICT r1: the round-1 parity tests stopped at 40-100 hops; here the named configurations run at their real
lengths, hard-mask threshold flips are counted bin by bin (tests/util.py:flip_aware_compare) and the allowance is
what was observed on the B200, not n_hops // 10.  Every run appends its figures to gpurun_out/long_parity.json
(copied to profiles/ per round)."""
import json
import os

import numpy as np
import pytest

from tests.util import flip_aware_compare, peak_norm_err
from zen_b200.synth import synth_audio

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_ABS, TOL_SNR = 1e-4, 80.0      # BASELINE.json north_star: max abs <= 1e-4, SNR >= 80 dB (peak-normalised)
# A flipped hard-mask bin must be a borderline decision of the oracle: |ratio - beta| / beta below this.  The margin
# a last-bit difference of the two FFTs can bridge grows as the bin gets weaker relative to the frame (the FFT error is
# absolute); over the 2583 to 10336-hop runs here the largest observed is 2.2e-5 (gpurun_out/long_parity.json).
MARGIN_TOL = 6e-5
FS = 44100


def _report(name, rec):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    p = os.path.join(d, "long_parity.json")
    try:
        cur = json.load(open(p))
    except Exception:  # noqa: BLE001
        cur = {}
    cur[name] = rec
    json.dump(cur, open(p, "w"), indent=1)


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a GPU")
    return t


@pytest.fixture(scope="module")
def zen():
    from zen_b200 import hps
    return hps


def _summ(r, n_hops):
    return {"hops": int(n_hops), "flip_hops": int(np.count_nonzero(r["flips"])), "flipped_bins": int(r["flips"].sum()),
            "worst_margin": float(r["worst_margin"]), "hops_checked": int(r["hops_checked"]),
            "max_abs_err": [float(e) for e in r["err"]], "snr_db": [float(s) for s in r["snr"]]}


def test_config4_full_length_streams(torch, zen, oracle):
    """configs[4]: 60 s real-time streams (2583 hops of 1024, percussive out, hard mask, copy-border), four of them:
    the per-hop path against the oracle on every hop, and the batched kernel equal to that path bit for bit"""
    from bench import fakert_hops
    hop, beta, n_streams = 1024, 2.5, 4
    n_hops = fakert_hops(60 * FS, hop)
    assert n_hops == 2583
    audio = np.stack([synth_audio(n_hops * hop, seed=1000 + s) for s in range(n_streams)])
    b = zen.HPRBatch(float(FS), hop, beta, zen.OUTPUT_PERCUSSIVE)
    batch_p = b.process(torch.from_numpy(audio).cuda())[1].cpu().numpy()
    b.close()
    recs = []
    for s in range(n_streams):
        o = oracle.OracleHPR(oracle.GEOM_GPU, float(FS), hop, beta, 2, oracle.CAUSAL, True)
        h = zen.HPR(float(FS), hop, beta, 2, 0, True)
        r = flip_aware_compare(h, o, audio[s], hop, 2, hard_mask=True, margin_tol=MARGIN_TOL)
        h.close()
        o.close()
        rec = _summ(r, n_hops)
        recs.append(rec)
        assert r["margin_ok"], ("mask mismatch on a non-borderline bin", r["worst_margin"])
        assert rec["flip_hops"] <= max(2, n_hops // 100), rec
        assert r["err"][1] <= TOL_ABS and r["snr"][1] >= TOL_SNR, rec
        assert np.array_equal(batch_p[s], r["got"][1]), "batched kernel != per-hop path on stream %d" % s
    _report("config4_rt1024_4x2583", recs)


def test_config3_sse_soft_full_length(torch, zen, oracle):
    """configs[3]: --soft-mask --sse at hop 512 on 60 s (5167 hops): no thresholds, so every hop must meet the tolerance"""
    from bench import fakert_hops
    hop, beta = 512, 2.5
    n_hops = fakert_hops(60 * FS, hop)
    assert n_hops == 5167
    audio = synth_audio(n_hops * hop, seed=4)
    o = oracle.OracleHPR(oracle.GEOM_GPU, float(FS), hop, beta, 2, oracle.CAUSAL, True)
    o.use_sse_filter()
    o.use_soft_mask()
    ref = o.run(audio, n_hops, want=(False, True, False))[1]
    b = zen.HPRBatch(float(FS), hop, beta, zen.OUTPUT_PERCUSSIVE, sse=True, soft=True)
    got = b.process(torch.from_numpy(audio[None]).cuda())[1][0].cpu().numpy()
    b.close()
    err, snr = peak_norm_err(got, ref)
    _report("config3_sse_soft_512_5167", {"hops": n_hops, "max_abs_err": err, "snr_db": snr})
    assert err <= TOL_ABS and snr >= TOL_SNR, (err, snr)


def test_config2_offline_two_pass_60s(torch, zen, oracle):
    """configs[2] (hop 4096 then 256, beta 2.5, hard mask) on 60 s through both passes.  Each pass is checked
    flip-aware at its full length on the streaming objects HPRIOffline drives (anticausal, hps.cu:38-48); the whole
    HPRIOffline::process output is then compared hop by hop, the hops that differ beyond the tolerance being exactly
    the ones a flipped bin explains (they must be few, and are counted)."""
    n = 60 * FS
    audio = synth_audio(n, seed=3)
    beta = 2.5
    rec = {}
    # pass 1 objects: H|P|R at hop 4096; pass 2: P at hop 256 (hps.cu:38-48)
    for name, hop, flags in (("pass1_hop4096", 4096, 7), ("pass2_hop256", 256, 2)):
        n_hops = -(-n // hop)
        a = np.zeros(n_hops * hop, np.float32)
        a[:n] = audio
        o = oracle.OracleHPR(oracle.GEOM_GPU, float(FS), hop, beta, flags, oracle.ANTICAUSAL, True)
        h = zen.HPR(float(FS), hop, beta, flags, 1, True)
        r = flip_aware_compare(h, o, a, hop, flags, hard_mask=True, margin_tol=MARGIN_TOL)
        h.close()
        o.close()
        rec[name] = _summ(r, n_hops)
        assert r["margin_ok"], (name, r["worst_margin"])
        # observed on the B200: 11 of 646 hops at hop 4096 (three 187-tap masks per hop), 30 of 10336 at hop 256
        assert rec[name]["flip_hops"] <= max(2, n_hops // 40), rec[name]
        for o_idx in range(3):
            if flags & (1 << o_idx):
                assert r["err"][o_idx] <= TOL_ABS and r["snr"][o_idx] >= TOL_SNR, (name, o_idx, rec[name])
    ref = oracle.offline_process(oracle.GEOM_GPU, float(FS), 4096, 256, beta, beta, audio)
    got = zen.HPRIOffline(float(FS), 4096, 256, beta, beta).process(audio)
    assert all(g.size == n for g in got) and not got[2].any() and not ref[2].any()
    whole = {}
    for nm, g, rf, hop in (("harmonic", got[0], ref[0], 4096), ("percussive", got[1], ref[1], 256)):
        pk = float(np.abs(rf).max())
        nh = n // hop
        e = np.abs(g[: nh * hop] - rf[: nh * hop]).reshape(nh, hop).max(axis=1) / pk
        bad = int(np.count_nonzero(e > TOL_ABS))
        whole[nm] = {"hops": nh, "hops_beyond_tol": bad, "max_abs_err_clean_hops": float(e[e <= TOL_ABS].max())}
        # a flipped bin in pass 1 changes the input of pass 2 for one 4096-hop (16 hops of 256) and its overlap
        assert bad <= max(4, nh // 50), (nm, whole[nm])
    rec["whole_process"] = whole
    _report("config2_offline_60s", rec)
