"""seeded inputs of the beat-tracker tests (shared by the golden generator and the tests); hops of 256 samples"""
import numpy as np

from zen_b200.synth import synth_audio


def make_input(fs, n_hops, kind, arg):
    n = n_hops * 256
    rng = np.random.default_rng(17)
    if kind == "clicks":      # a decaying noise burst every 60 / arg seconds on a quiet tone: a drum track at `arg` BPM
        x = (0.02 * np.sin(2 * np.pi * 220.0 * np.arange(n) / fs)).astype(np.float64)
        period = 60.0 / arg * fs
        blen = int(0.03 * fs)
        env = np.exp(-np.arange(blen) / (0.006 * fs))
        pos = 0.1 * fs
        while pos + blen < n:
            i0 = int(pos)
            x[i0:i0 + blen] += 0.7 * env * rng.standard_normal(blen)
            pos += period
        return np.clip(x, -1, 1).astype(np.float32)
    if kind == "synth":
        return synth_audio(n, seed=arg, fs=int(fs))
    if kind == "noise":
        return (0.1 * np.random.default_rng(arg).standard_normal(n)).astype(np.float32)
    if kind == "silence":
        return np.zeros(n, np.float32)
    raise ValueError(kind)


CASES = [
    ("clicks120_44k", 44100.0, 3000, "clicks", 120.0),
    ("clicks96_44k", 44100.0, 3000, "clicks", 96.0),
    ("clicks150_48k", 48000.0, 3000, "clicks", 150.0),
    ("synth_44k", 44100.0, 2500, "synth", 3),
    ("noise_44k", 44100.0, 1500, "noise", 4),
    ("silence_44k", 44100.0, 600, "silence", 0),
]
