"""`zen offline | fakert` command line (reference: zen/main.cu:20-63, zen/offline.h, zen/fakert.h): flags,
defaults, output files, peak normalisation and PCM16 encoding."""
import os
import struct
import subprocess

import numpy as np
import pytest

from zen_b200.synth import synth_audio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZEN = os.path.join(ROOT, "zen_b200", "bin", "zen")


def write_wav(path, x, fs=44100, channels=1):
    q = np.clip(np.round(x * 32767.0), -32768, 32767).astype("<i2")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + q.nbytes) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, channels, fs, fs * 2 * channels, 2 * channels, 16))
        f.write(b"data" + struct.pack("<I", q.nbytes) + q.tobytes())
    return (q.astype(np.float32) / np.float32(32767.0)).astype(np.float32)


def read_wav(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    fmt, ch, fs = struct.unpack("<HHI", b[20:28])
    n = struct.unpack("<I", b[40:44])[0]
    return fs, ch, np.frombuffer(b[44:44 + n], dtype="<i2").astype(np.int32)


def run(*args):
    return subprocess.run([ZEN] + list(args), capture_output=True, text=True, timeout=300)


def test_version_help_and_errors(tmp_path):
    assert os.path.exists(ZEN), "build with make -C zen_b200/csrc"
    assert run("version").stdout == "version 1.0\n"
    assert run("-v").stdout == "version 1.0\n"
    assert "zen offline" in run("help").stdout and "zen fakert" in run("--help").stdout
    assert run("bogus").returncode != 0
    assert run("offline").returncode != 0                      # -i is required
    wav = str(tmp_path / "a.wav")
    write_wav(wav, synth_audio(8192, seed=1))
    r = run("offline", "-i", wav, "--hps", "--cpu")
    assert r.returncode == 2 and "GPU path only" in r.stderr   # no CPU backend in this build
    lfs = str(tmp_path / "lfs.wav")
    open(lfs, "w").write("version https://git-lfs.github.com/spec/v1\noid sha256:4e2d\nsize 323186\n")
    r = run("fakert", "-i", lfs, "--hps")
    assert r.returncode == 1 and "RIFF" in r.stderr


@pytest.mark.gpu
def test_offline_cli_matches_library(tmp_path):
    from zen_b200 import hps
    x = synth_audio(10 * 4096 + 11, seed=21)
    stereo = np.stack([x, x], axis=1).reshape(-1)              # identical channels: the mono fold returns x
    wav = str(tmp_path / "in.wav")
    xq = write_wav(wav, stereo, channels=2)[0::2]
    prefix = str(tmp_path / "out")
    r = run("offline", "-i", wav, "--hps", "4096", "2.5", "256", "2.5", "-o", prefix)
    assert r.returncode == 0, r.stderr
    assert "GPU/CUDA/thrust: 2-pass HPR-I-Offline took" in r.stdout and "with HPR-I separation using harmonic params: 4096,2.5" in r.stdout
    ref = hps.HPRIOffline(44100.0, 4096, 256, 2.5, 2.5).process(xq)
    for name, o in (("_harm.wav", ref[0]), ("_perc.wav", ref[1])):
        fs, ch, got = read_wav(prefix + name)
        assert (fs, ch, got.size) == (44100, 1, xq.size)
        exp = np.round(o / np.abs(o).max() * 32767.0).astype(np.int32)
        assert np.abs(got - exp).max() <= 1
    assert os.path.exists(prefix + "_residual.wav")
    # --only-percussive writes a single file; defaults are 4096 / 2.0 / 256 / 2.0
    prefix2 = str(tmp_path / "p")
    r = run("offline", "-i", wav, "--hps", "-o", prefix2, "--only-percussive", "--soft-mask")
    assert r.returncode == 0 and "harmonic params: 4096,2, percussive params: 256,2" in r.stdout
    assert os.path.exists(prefix2 + "_perc.wav") and not os.path.exists(prefix2 + "_harm.wav")
    # invalid hop pair -> ZgException -> non-zero exit
    assert run("offline", "-i", wav, "--hps", "4096", "2.0", "300", "2.0").returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
def test_fakert_cli(tmp_path, oracle, resident):
    hop = 1024
    x = synth_audio(30 * hop + 500, seed=5)
    wav, out = str(tmp_path / "in.wav"), str(tmp_path / "perc.wav")
    xq = write_wav(wav, x)
    env = dict(os.environ)
    env["ZEN_RESIDENT"] = "1" if resident else "0"
    r = subprocess.run([ZEN, "fakert", "-i", wav, "--hps", "1024", "2.5", "-o", out], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    n_chunks = oracle.fakert_n_chunks(xq.size, hop)
    assert "into %d chunks of size %d" % (n_chunks, hop) in r.stdout
    assert "PRealtime GPU:  Δn = 1024" in r.stdout and "average processing duration(us)" in r.stdout
    o = oracle.OracleHPR(oracle.GEOM_GPU, 44100.0, hop, 2.5, oracle.OUT_P, oracle.CAUSAL, True)
    perc = o.run(xq, n_chunks, want=(False, True, False))[1]
    exp = xq.copy()                      # the unprocessed tail stays raw input (fakert.h:132)
    exp[: n_chunks * hop] = perc
    exp = exp / max(-exp.min(), exp.max())
    fs, ch, got = read_wav(out)
    assert got.size == xq.size
    err = np.abs(got - np.round(exp * 32767.0)).astype(np.int64)
    assert np.mean(err > 1) < 0.02 and err.max() <= 40         # 1 LSB everywhere but hops touched by a hard-mask threshold flip
