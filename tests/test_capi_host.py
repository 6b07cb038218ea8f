"""The C-ABI library loads, exports every symbol include/zen_b200.h declares, and
its host-side logic (derived sizes, window, argument / geometry errors) matches
the oracle — no kernel is launched, so this runs on a box without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from zen_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zen_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zen_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(n for n in _lib.EXPORTS) == names


def test_version_and_device_count():
    L = _lib.lib()
    assert b"sm_100a" in L.zen_b200_version()
    assert L.zen_device_count() >= 0


@pytest.mark.parametrize("fs,hop,causal", [(44100.0, 256, 1), (44100.0, 512, 0), (44100.0, 1024, 1), (44100.0, 2048, 0),
                                           (44100.0, 4096, 1), (48000.0, 256, 0), (48000.0, 1024, 1), (22050.0, 512, 1)])
def test_geometry_matches_oracle(oracle, fs, hop, causal):
    g = _lib.ZenGeometry()
    assert _lib.lib().zen_hpr_geometry(fs, hop, causal, ctypes.byref(g)) == _lib.ZEN_OK
    o = oracle.geometry(fs, hop, bool(causal))
    assert (g.hop, g.nwin, g.nfft, g.l_harm, g.l_perc, g.lag, g.stft_width) == (o.hop, o.nwin, o.nfft, o.l_harm, o.l_perc, o.lag, o.stft_width)
    assert np.float32(g.cola_factor) == np.float32(o.cola)


@pytest.mark.parametrize("n", [512, 1024, 2048, 8192])
def test_window_bit_identical_to_oracle(oracle, n):
    for sqrt in (True, False):
        w = np.zeros(n, dtype=np.float32)
        assert _lib.lib().zen_window(0 if sqrt else 1, n, w.ctypes.data) == _lib.ZEN_OK
        assert np.array_equal(w, oracle.window(n, sqrt=sqrt))


def test_argument_and_geometry_errors_without_compute():
    L = _lib.lib()
    # ZgException cases are decided on the host before anything is launched
    assert L.zen_median_filter(9, 9, 10, _lib.TIME_CAUSAL, 0, 1, 1, None) == _lib.ZEN_ERR_GEOMETRY
    assert L.zen_median_filter(9, 9, 10, _lib.FREQUENCY, 1, 1, 1, None) == _lib.ZEN_ERR_GEOMETRY
    assert L.zen_box_filter(9, 9, 10, _lib.TIME_ANTICAUSAL, 1, 1, None) == _lib.ZEN_ERR_GEOMETRY
    assert L.zen_median_filter(9, 9, 3, 0, 0, None, None, None) == _lib.ZEN_ERR_ARG
    assert L.zen_fft_c2c(1000, 1, 0, None) == _lib.ZEN_ERR_UNSUPPORTED
    assert L.zen_fft_c2c(1 << 17, 1, 0, None) == _lib.ZEN_ERR_UNSUPPORTED
    a = np.zeros(8, np.float32)
    assert L.zen_offline_process(44100.0, 4096, 300, 2.0, 2.0, 0, a.ctypes.data, 8, a.ctypes.data, a.ctypes.data,
                                 a.ctypes.data) == _lib.ZEN_ERR_GEOMETRY     # hps.cu:33-36


def test_no_cpu_fallback():
    """without a CUDA device every computing entry point must fail loudly"""
    L = _lib.lib()
    if L.zen_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert L.zen_hpr_create(ctypes.byref(h), 44100.0, 1024, 2.5, 7, 0, 1) == _lib.ZEN_ERR_CUDA
    b = ctypes.c_void_p()
    assert L.zen_hpr_batch_create(ctypes.byref(b), 44100.0, 1024, 2.5, 2, 0, 0, 1, 1) == _lib.ZEN_ERR_CUDA
    a = np.zeros(4096 * 4, np.float32)
    assert L.zen_offline_process(44100.0, 4096, 256, 2.0, 2.0, 0, a.ctypes.data, a.size, a.ctypes.data, a.ctypes.data,
                                 a.ctypes.data) == _lib.ZEN_ERR_CUDA
    from zen_b200 import hps
    with pytest.raises(_lib.ZenCudaError):
        hps.HPRRealtime(44100.0, 1024, 2.5, hps.OUTPUT_PERCUSSIVE)


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/"""
    banned = re.compile(r"(import\s+oracle|from\s+oracle|oraclebind|refbind|libzen_oracle|libzen_ref|hpr_oracle\.h|#include\s+[<\"][^>\"]*oracle)")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zen_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not banned.search(txt), os.path.join(dirpath, f)


@pytest.mark.parametrize("hop", [32, 64, 128, 256, 512, 1024, 2048, 4096, 1, 2, 3, 4, 5, 11, 12, 13, 1000])
def test_tagged_group_format_round_trip(hop):
    """the resident session's staging format {x[3g], x[3g+1], x[3g+2], tag} (hpr_launch.cuh RtCtrl): packing then
    unpacking is the identity for every hop length (vector body, scalar tail, ragged last group), every group carries
    the tag (stored XOR a hash of its samples), the padding of the last group is zero, and a single stale or
    half-written group stops the unpack exactly there"""
    import ctypes
    from zen_b200 import _lib
    L = _lib.lib()
    L.zen_rt_pack_groups.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p]
    L.zen_rt_unpack_groups.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p]
    rng = np.random.default_rng(hop)
    x = rng.standard_normal(hop).astype(np.float32)
    x[rng.integers(0, hop)] = -0.0
    groups = (hop + 2) // 3
    raw = np.zeros(groups * 4 + 4, np.uint32)
    off = (-raw.ctypes.data // 4) % 4                       # 16-byte aligned view
    g = raw[off:off + groups * 4]
    assert g.ctypes.data % 16 == 0
    tag = (123456 << 8) | 0x61
    assert L.zen_rt_pack_groups(x.ctypes.data, hop, tag, g.ctypes.data) == 0
    gv = g.reshape(groups, 4)

    def rotl(v, n):
        return ((v << np.uint32(n)) | (v >> np.uint32(32 - n))).astype(np.uint32)
    # The tag words are stored XOR a hash of the samples (zen_group_key, hpr_core.cuh): the four groups of a whole
    # 64-byte line share the line hash V[l] = y[l] ^ rotl(y[4+l], 11) ^ rotl(y[8+l], 22) of their twelve samples, the
    # groups behind the last whole line use zen_group_hash(x0, x1, x2).  Either way a group validates itself.
    xu = np.zeros(groups * 3, np.uint32)
    xu[:hop] = x.view(np.uint32)
    lg = 4 * (hop // 12)
    key = np.zeros(groups, np.uint32)
    for b in range(lg // 4):
        y = xu[12 * b:12 * b + 12]
        key[4 * b:4 * b + 4] = y[0:4] ^ rotl(y[4:8], 11) ^ rotl(y[8:12], 22)
    t = xu[3 * lg:].reshape(-1, 3)
    key[lg:] = t[:, 0] ^ rotl(t[:, 1], 11) ^ rotl(t[:, 2], 22)
    assert np.all((gv[:, 3] ^ key) == tag)
    flat = gv[:, :3].reshape(-1)
    assert np.array_equal(flat[:hop], x.view(np.uint32))    # bit patterns, -0.0 included
    assert np.all(flat[hop:] == 0)
    y = np.full(hop + 4, 7.0, np.float32)
    assert L.zen_rt_unpack_groups(g.ctypes.data, hop, tag, y.ctypes.data) == groups
    assert np.array_equal(y[:hop].view(np.uint32), x.view(np.uint32))
    assert np.all(y[hop:] == 7.0)                           # nothing written past the hop
    # one stale group: the unpack stops there and leaves the rest alone
    for stale in sorted({0, groups // 2, groups - 1}):
        buf = np.zeros(groups * 4 + 4, np.uint32)
        o2 = (-buf.ctypes.data // 4) % 4
        g2 = buf[o2:o2 + groups * 4]
        g2[:] = g
        if stale % 2 == 0:
            g2[4 * stale + 3] ^= np.uint32((tag - 256) ^ tag)  # the previous request's tag on the same samples
        else:
            g2[4 * stale + 1] ^= np.uint32(1 << 7)          # a group caught half-written: a sample that does not belong to the tag word
        y2 = np.full(hop + 4, 7.0, np.float32)
        stop = 4 * (stale // 4) if stale < lg else stale     # a whole line is taken or left
        assert L.zen_rt_unpack_groups(g2.ctypes.data, hop, tag, y2.ctypes.data) == stop
        assert np.array_equal(y2[:3 * stop].view(np.uint32), x[:3 * stop].view(np.uint32))
        assert np.all(y2[3 * stop:] == 7.0)                 # nothing of the stale line / group or behind it was written
    assert L.zen_rt_pack_groups(x.ctypes.data, hop, tag, g.ctypes.data + 4) != 0   # misaligned staging buffer


@pytest.mark.parametrize("nfft", [128, 256, 512, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_cluster_split_owns_every_bin_exactly_once(nfft, cluster):
    """hpr_split_ranges (hpr_core.cuh): the CTAs of the resident cluster kernel divide the nfft/2 + 1 half-spectrum bins
    by pairs (k, M - k); every bin must have exactly one owner, a CTA's two ranges must not overlap, the pair ranges
    must tile [0, M/2], and the ranges must be the mirror images of the pairs"""
    import ctypes
    from zen_b200 import _lib
    L = _lib.lib()
    M = nfft // 2
    owners = np.zeros(M + 1, np.int32)
    next_k = 0
    for rank in range(cluster):
        r = (ctypes.c_int * 6)()
        assert L.zen_rt_split_ranges(nfft, rank, cluster, r) == 0
        k0, k1, a0, a1, b0, b1 = list(r)
        assert k0 == next_k and k0 <= k1 <= M // 2 + 1
        next_k = k1
        assert (a0, a1) == (k0, k1) or k0 == k1
        assert a1 <= b0 or k0 == k1                            # the two ranges of one CTA do not overlap
        owners[a0:a1] += 1
        owners[b0:b1] += 1
        own = set(range(a0, a1)) | set(range(b0, b1))
        assert own == {k for k in range(k0, k1)} | {M - k for k in range(k0, k1)}
    assert next_k == M // 2 + 1
    assert np.all(owners == 1)
    assert L.zen_rt_split_ranges(nfft, cluster, cluster, (ctypes.c_int * 6)()) != 0


def test_portable_host_path_of_the_tagged_groups():
    """the host code of the resident session no longer depends on x86 SSE intrinsics - a portable path
    (GCC vector extensions, one aligned 16-byte access per group) is selected off x86-64, or by -DZEN_PORTABLE_GROUPS.
    Built here with it forced on and put through the very same format tests as the SSE2 path."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "zen_b200", "csrc"), "portable"])
    so = os.path.join(root, "tests", "cpp", "_build", "libzen_b200_portable.so")
    assert os.path.exists(so)
    env = dict(os.environ, ZEN_B200_LIB=so)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", os.path.join(root, "tests", "test_capi_host.py"), "-k",
                        "tagged_group_format_round_trip or exports"], capture_output=True, text=True, env=env, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
