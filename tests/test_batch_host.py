"""Host-buffer entry points of the batched path (zen_hpr_batch_process_host, ..._pcm16) and the argument handling of
HPRBatch: staging buffers regrow with the request, the soft / SSE residual is zeros as N process_next_hop calls of the
reference emit (libzen/hps.cu:435-449, 562), strided inputs, PCM16 on both sides of the link
(zen/offline.h:88-117, 180-223; vendor/libnyquist/src/Common.cpp:332-337)."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import np_model

FS, HOP, BETA = 44100.0, 1024, 2.5


def test_div32767_exact():
    """csrc/pcm.cu evaluates (float)s / 32767.f as q = s*r, q' = fma(fma(-q, 32767, s), r, q) with r = RN(1/32767).
    Restated here in exact rational arithmetic with one rounding per operation: equal to the IEEE quotient for every
    int16 s."""
    f32 = np.float32

    def rn(fr):
        if fr == 0:
            return f32(0)
        x = f32(float(fr))
        cands = [x, np.nextafter(x, f32(np.inf)), np.nextafter(x, f32(-np.inf))]
        return min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(c.view(np.uint32)) & 1))

    r = rn(Fraction(1, 32767))
    assert r == f32(1.0) / f32(32767.0)
    rf = Fraction(float(r))
    for s in range(-32768, 32768):
        q0 = Fraction(float(rn(Fraction(s) * rf)))
        rem = Fraction(float(rn(Fraction(s) - q0 * 32767)))
        q = rn(q0 + rem * rf)
        assert q == f32(s) / f32(32767.0), s


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a GPU")
    return t


def _audio(n_streams, n, seed=0):
    from zen_b200.synth import synth_audio
    return np.stack([synth_audio(n, seed=seed + s) for s in range(n_streams)])


@pytest.mark.gpu
def test_process_host_regrows_its_staging_buffers(torch):
    """one batch object, three calls - longer rows, then more outputs, then more streams - each
    equal to the device-resident entry point"""
    from zen_b200 import hps
    b = hps.HPRBatch(FS, HOP, BETA, 7)
    ref = hps.HPRBatch(FS, HOP, BETA, 7)
    for n_streams, n_hops, want in ((3, 8, (False, True, False)), (3, 40, (False, True, False)), (2, 40, (True, True, True)),
                                    (7, 64, (True, True, True)), (2, 5, (True, False, True))):
        x = _audio(n_streams, n_hops * HOP, seed=n_hops)
        outs = [np.full_like(x, np.nan) if w else None for w in want]
        b.process_host(x, outs)
        exp = ref.process(torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        for o in range(3):
            if want[o]:
                assert np.array_equal(outs[o], exp[o].cpu().numpy()), (n_streams, n_hops, o)
    b.close()
    ref.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["soft", "sse"])
def test_batched_residual_of_soft_and_sse_is_zero(torch, variant, oracle):
    """residual_out is only rotated and zero-filled when the mask is soft / SSE, so the batch
    emits zeros - into caller-provided memory too, on both entry points"""
    from zen_b200 import hps
    n_streams, n_hops = 2, 12
    x = _audio(n_streams, n_hops * HOP, seed=5)
    b = hps.HPRBatch(FS, HOP, BETA, 7, soft=variant == "soft", sse=variant == "sse")
    xd = torch.from_numpy(x).cuda()
    outs = [torch.full_like(xd, float("nan")) for _ in range(3)]
    b.process(xd, outs)
    torch.cuda.synchronize()
    assert not outs[2].cpu().numpy().any()
    houts = [np.full_like(x, np.nan) for _ in range(3)]
    b.process_host(x, houts)
    assert not houts[2].any()
    for o in range(2):
        assert np.array_equal(houts[o], outs[o].cpu().numpy())
    # and that is what the oracle's N process_next_hop calls emit
    oh = oracle.OracleHPR(oracle.GEOM_GPU, FS, HOP, BETA, 7, oracle.CAUSAL, True)
    (oh.use_soft_mask if variant == "soft" else oh.use_sse_filter)()
    assert not oh.run(x[0], n_hops)[2].any()
    b.close()


@pytest.mark.gpu
def test_batch_strides_and_argument_checks(torch):
    """a sliced input (row stride wider than the row) with freshly allocated outputs; misuse raises"""
    from zen_b200 import hps
    n_hops = 10
    x = torch.from_numpy(_audio(3, (n_hops + 6) * HOP, seed=9)).cuda()
    view = x[:, : n_hops * HOP]
    b = hps.HPRBatch(FS, HOP, BETA, hps.OUTPUT_PERCUSSIVE)
    got = b.process(view)[1]
    exp = b.process(view.contiguous())[1]
    torch.cuda.synchronize()
    assert got.shape == view.shape and torch.equal(got, exp)
    wide = torch.empty_like(x)
    got2 = b.process(view, [None, wide[:, : n_hops * HOP], None])[1]
    torch.cuda.synchronize()
    assert torch.equal(got2, exp)
    with pytest.raises(ValueError):
        b.process(x[:, : n_hops * HOP + 3])            # not a multiple of hop
    with pytest.raises(ValueError):
        b.process(view.t().contiguous().t())            # rows not contiguous
    with pytest.raises(ValueError):
        b.process(view, [None, torch.empty((3, 5), device="cuda"), None])
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("flags,n_streams,n_hops", [(2, 5, 30), (7, 2, 17)])
def test_process_host_pcm16_equals_decode_process_encode(torch, flags, n_streams, n_hops):
    """PCM16 in, peak-normalised PCM16 out: bit-identical to the libnyquist / command-line host code restated in
    oracle/np_model.py applied around the float entry point"""
    from zen_b200 import hps
    rng = np.random.default_rng(flags)
    x = _audio(n_streams, n_hops * HOP, seed=31)
    pcm = np.round(x * 32767).astype(np.int16)
    pcm[0, 0] = -32768
    if n_streams > 2:
        pcm[2] = 0                                        # a silent stream stays silent
    b = hps.HPRBatch(FS, HOP, BETA, flags)
    dec = np.stack([np_model.pcm16_decode_mono(pcm[s], 1) for s in range(n_streams)])
    exp = b.process(torch.from_numpy(dec).cuda())
    torch.cuda.synchronize()
    outs = [np.full(pcm.shape, 12345, np.int16) if flags & (1 << o) else None for o in range(3)]
    peaks = [np.full(n_streams, np.nan, np.float32) if flags & (1 << o) else None for o in range(3)]
    b.process_host_pcm16(pcm, outs, peaks)
    for o in range(3):
        if not flags & (1 << o):
            continue
        e = exp[o].cpu().numpy()
        for s in range(n_streams):
            q, pk = np_model.pcm16_encode_normalized(e[s])
            assert peaks[o][s] == pk, (o, s)
            assert np.array_equal(outs[o][s], q), (o, s)
    del rng
    b.close()
