"""The index arithmetic the CUDA kernels use (half spectrum, one consumed row,
mirrored windows, --nocopybord effective masks) reproduces the oracle's
full-matrix restatement of the reference.  CPU only."""
import numpy as np
import pytest

from tests.halfspec_model import HalfSpecHPR
from tests.util import peak_norm_err
from zen_b200.synth import synth_audio

CASES = [
    # fs, hop, beta, flags, causal, copy_bord, sse, soft, n_hops
    (44100.0, 1024, 2.5, 7, True, True, False, False, 14),
    (44100.0, 1024, 2.5, 7, True, False, False, False, 14),
    (48000.0, 256, 2.0, 7, True, True, False, False, 40),
    (48000.0, 256, 2.0, 7, True, False, False, False, 40),
    (48000.0, 256, 2.0, 7, False, True, False, False, 40),
    (48000.0, 256, 2.0, 7, False, False, False, False, 40),
    (44100.0, 512, 2.5, 7, True, True, True, True, 24),
    (44100.0, 512, 2.5, 3, True, True, False, True, 24),
    (44100.0, 2048, 2.5, 7, True, False, False, False, 6),
    (44100.0, 2048, 2.5, 7, False, False, False, False, 6),
    (44100.0, 4096, 2.5, 6, False, True, False, False, 5),
]


@pytest.mark.parametrize("fs,hop,beta,flags,causal,cb,sse,soft,n_hops", CASES)
def test_halfspec_matches_oracle(oracle, fs, hop, beta, flags, causal, cb, sse, soft, n_hops):
    audio = synth_audio(n_hops * hop, seed=hop + int(cb))
    o = oracle.OracleHPR(oracle.GEOM_GPU, fs, hop, beta, flags, oracle.CAUSAL if causal else oracle.ANTICAUSAL, cb)
    if sse:
        o.use_sse_filter()
    if soft:
        o.use_soft_mask()
    ref = o.run(audio)
    geom = dict(stft_width=o.stft_width, lag=o.lag, l_harm=o.l_harm, l_perc=o.l_perc)
    m = HalfSpecHPR(fs, hop, beta, flags, causal, cb, geom, o.get("window"), o.cola, sse=sse, soft_mask=soft)
    got = m.run(audio)
    for name, a, b in zip("HPR", got, ref):
        err, snr = peak_norm_err(a, b)
        assert err <= 1e-4 and snr >= 80.0, (name, err, snr)
