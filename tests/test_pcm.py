"""The on-disk format either side of the path: PCM16 decode + mono fold and peak
normalisation + PCM16 encode, batched on the device, against the NumPy restatement of libnyquist's / the zen command
line's host code (oracle/np_model.py)."""
import os

import numpy as np
import pytest

from oracle import np_model


def test_pcm16_oracle_known_answers():
    """the conventions of vendor/libnyquist/include/libnyquist/Common.h:296-302, 669-675 and src/Common.cpp:332-337"""
    s = np.array([0, 1, -1, 32767, -32767, -32768, 16384], np.int16)
    f = np_model.pcm16_decode_mono(s, 1)
    assert f.dtype == np.float32
    assert f[3] == np.float32(1.0) and f[4] == np.float32(-1.0)
    assert f[5] == np.float32(-32768.0) / np.float32(32767.0) and f[5] < -1.0      # int16 min decodes below -1
    assert f[1] == np.float32(1.0) / np.float32(32767.0)                             # a division, not a multiplication by 1/32767
    st = np.array([100, 300, -32768, -32768, 32767, -32767], np.int16)
    m = np_model.pcm16_decode_mono(st, 2)
    assert np.array_equal(m, ((st[0::2].astype(np.float32) / np.float32(32767)) + (st[1::2].astype(np.float32) / np.float32(32767))) / np.float32(2))
    assert m[2] == 0.0
    # lroundf: halfway cases away from zero
    assert list(np_model.lroundf(np.array([0.5, -0.5, 1.5, 2.5, -2.5, 0.49999997], np.float32))) == [1, -1, 2, 3, -3, 0]
    # encode: the peak maps to +-32767 whatever its sign
    q, peak = np_model.pcm16_encode_normalized(np.array([0.25, -0.5, 0.1], np.float32))
    assert peak == np.float32(0.5) and list(q) == [16384, -32767, 6553]
    # decoding PCM16 and encoding it again is the identity when the file reaches full scale
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32767, 32768, 5000).astype(np.int16)
    pcm[17] = 32767
    q, peak = np_model.pcm16_encode_normalized(np_model.pcm16_decode_mono(pcm, 1))
    assert peak == np.float32(1.0) and np.array_equal(q, pcm)
    # a silent signal stays silent (the reference divides by zero there)
    q, peak = np_model.pcm16_encode_normalized(np.zeros(8, np.float32))
    assert peak == 0 and not q.any()


def test_pcm16_oracle_matches_the_wav_helpers_of_the_cli_tests(tmp_path):
    """tests/test_cli.py writes / reads real wav files with the same conventions (PCM16, x * 32767 rounded)"""
    from tests.test_cli import read_wav, write_wav
    rng = np.random.default_rng(1)
    x = (0.8 * rng.standard_normal(3000)).clip(-1, 1).astype(np.float32)
    xq = write_wav(str(tmp_path / "a.wav"), x)
    fs, ch, got = read_wav(str(tmp_path / "a.wav"))
    assert fs == 44100 and ch == 1
    assert np.array_equal(np_model.pcm16_decode_mono(got.astype(np.int16), 1), xq)


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nyq_pcm.npz")


def test_pcm16_oracle_pinned_by_libnyquist():
    """oracle/np_model.py against the UNMODIFIED libnyquist conversion code compiled from /root/reference/vendor
    (oracle/Makefile ref_nyq, fixtures by oracle/ref/make_pcm_golden.py): every PCM16 code decoded, the stereo fold,
    the PCM16 encode on halfway cases, and the command line's peak normalisation (zen/offline.h:180-192)."""
    g = np.load(GOLD)
    all16 = np.arange(-32768, 32768, dtype=np.int16)
    assert np.array_equal(np_model.pcm16_decode_mono(all16, 1).view(np.uint32), g["decode_all_int16"].view(np.uint32))
    assert np.array_equal(np_model.pcm16_decode_mono(g["stereo_pcm"], 2).view(np.uint32), g["stereo_mono"].view(np.uint32))
    e = g["encode_in"]
    assert np.array_equal(np_model.lroundf((e * np.float32(32767.0)).astype(np.float32)).astype(np.int16), g["encode_out"])
    for i in range(5):
        q, pk = np_model.pcm16_encode_normalized(g["norm%d_in" % i])
        assert pk == g["norm%d_peak" % i] and np.array_equal(q, g["norm%d_out" % i]), i
    # silence: 0 / 0 = NaN in the reference, which its x86-64 build converts to 0 - the same PCM16 our path writes
    assert not g["silence_out"].any()
    q, pk = np_model.pcm16_encode_normalized(np.zeros(64, np.float32))
    assert np.array_equal(q, g["silence_out"])


@pytest.mark.gpu
def test_pcm16_kernels_pinned_by_libnyquist():
    """csrc/pcm.cu against the same libnyquist fixtures, bit for bit"""
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a GPU")
    from zen_b200 import hps
    g = np.load(GOLD)
    all16 = np.arange(-32768, 32768, dtype=np.int16)
    got = hps.pcm16_decode_mono(t.from_numpy(all16[None]).cuda(), 1).cpu().numpy()[0]
    assert np.array_equal(got.view(np.uint32), g["decode_all_int16"].view(np.uint32))
    got = hps.pcm16_decode_mono(t.from_numpy(g["stereo_pcm"][None]).cuda(), 2).cpu().numpy()[0]
    assert np.array_equal(got.view(np.uint32), g["stereo_mono"].view(np.uint32))
    x = np.stack([g["norm%d_in" % i] for i in range(5)] + [np.zeros(g["norm0_in"].size, np.float32)])
    q, pk = hps.pcm16_encode_normalized(t.from_numpy(x).cuda())
    q, pk = q.cpu().numpy(), pk.cpu().numpy()
    for i in range(5):
        assert pk[i] == g["norm%d_peak" % i] and np.array_equal(q[i], g["norm%d_out" % i]), i
    assert not q[5].any()


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a GPU")
    return t


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams,n_frames,channels", [(1, 1, 1), (3, 1000, 1), (2, 4097, 2), (5, 161571, 1), (64, 44100, 2), (1, 0, 1)])
def test_pcm16_decode_bit_exact(torch, n_streams, n_frames, channels):
    from zen_b200 import hps
    rng = np.random.default_rng(n_frames + channels)
    pcm = rng.integers(-32768, 32768, (n_streams, n_frames * channels)).astype(np.int16)
    if n_frames:
        pcm[0, :channels] = -32768
    got = hps.pcm16_decode_mono(torch.from_numpy(pcm).cuda(), channels).cpu().numpy()
    assert got.shape == (n_streams, n_frames)
    for s in range(n_streams):
        assert np.array_equal(got[s].view(np.uint32), np_model.pcm16_decode_mono(pcm[s], channels).view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams,n", [(1, 1), (3, 1000), (4, 161571), (96, 50001), (2, 0)])
def test_pcm16_encode_normalized_bit_exact(torch, n_streams, n):
    from zen_b200 import hps
    rng = np.random.default_rng(n + n_streams)
    x = (rng.standard_normal((n_streams, n)) * rng.uniform(1e-3, 30.0, (n_streams, 1))).astype(np.float32)
    if n_streams > 1 and n:
        x[1] = 0.0                                      # a silent stream
        x[2, n // 2] = -np.abs(x[2]).max() * 2          # the peak is a negative sample
    q, peaks = hps.pcm16_encode_normalized(torch.from_numpy(x).cuda())
    q, peaks = q.cpu().numpy(), peaks.cpu().numpy()
    for s in range(n_streams):
        eq, ep = np_model.pcm16_encode_normalized(x[s])
        assert peaks[s] == ep
        assert np.array_equal(q[s], eq)


@pytest.mark.gpu
def test_pcm16_round_trip_and_strides(torch):
    """full-scale PCM16 survives decode -> encode unchanged; rows may sit in a wider buffer (strides in elements)"""
    from zen_b200 import hps
    rng = np.random.default_rng(7)
    big = rng.integers(-32767, 32768, (6, 30000)).astype(np.int16)
    big[:, 5] = 32767
    d = torch.from_numpy(big).cuda()
    view = d[:, :20000]                                  # stride 30000, 20000 frames
    f = hps.pcm16_decode_mono(view, 1)
    q, peaks = hps.pcm16_encode_normalized(f)
    assert np.all(peaks.cpu().numpy() == 1.0)
    assert np.array_equal(q.cpu().numpy(), big[:, :20000])


@pytest.mark.gpu
def test_pcm16_encode_fast_division_is_the_ieee_one(torch):
    """The encode kernel divides by the row's peak with a hoisted reciprocal and two residual corrections instead of one
    IEEE division per sample (csrc/pcm.cu, pcm_from_float_fast).  Peaks with awkward significands (all ones, one ulp
    around powers of two, random), samples drawn uniformly and packed around the PCM16 rounding boundaries
    (k + 1/2) / 32767 x peak +- a few ulps: every integer must be the one float32 division gives."""
    from zen_b200 import hps
    rng = np.random.default_rng(99)
    n_rows, n = 96, 1 << 19
    mant = rng.integers(0, 1 << 23, n_rows).astype(np.uint32)
    mant[:8] = [0x7fffff, 0x7ffffe, 0, 1, 0x400000, 0x3fffff, 0x555555, 0x2aaaaa]
    expo = rng.integers(127 - 40, 127 + 40, n_rows).astype(np.uint32)
    expo[8:16] = 127
    expo[16:20] = [127 + 70, 127 - 70, 127 + 100, 127 - 100]      # outside the fast path's range: IEEE division per sample
    peaks = ((expo << 23) | mant).view(np.float32)
    x = np.empty((n_rows, n), np.float32)
    for r in range(n_rows):
        pk = peaks[r]
        u = rng.uniform(-1.0, 1.0, n // 2).astype(np.float32) * pk
        k = rng.integers(-32767, 32767, n - n // 2 - 1).astype(np.float64)
        b = ((k + 0.5) / 32767.0 * np.float64(pk)).astype(np.float32)
        b = (b.view(np.int32) + rng.integers(-3, 4, b.size).astype(np.int32)).view(np.float32)
        row = np.concatenate([u, b, np.array([pk], np.float32)])
        np.clip(row, -pk, pk, out=row)
        row[-1] = pk if r % 2 else -pk
        x[r] = row
    q, pk_out = hps.pcm16_encode_normalized(torch.from_numpy(x).cuda())
    q, pk_out = q.cpu().numpy(), pk_out.cpu().numpy()
    assert np.array_equal(pk_out.view(np.uint32), peaks.view(np.uint32))
    bad = 0
    for r in range(n_rows):
        eq, _ = np_model.pcm16_encode_normalized(x[r])
        bad += int((eq != q[r]).sum())
    assert bad == 0
