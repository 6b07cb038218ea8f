"""seeded inputs of the McLeod pitch tests (shared by the golden generator and the tests)"""
import numpy as np

from zen_b200.synth import synth_audio


def make_input(n, fs, kind, arg):
    t = np.arange(n) / fs
    if kind == "tone":
        return (0.6 * np.sin(2 * np.pi * arg * t) + 0.2 * np.sin(2 * np.pi * 2 * arg * t) + 0.1 * np.sin(2 * np.pi * 3 * arg * t)).astype(np.float32)
    if kind == "synth":
        return synth_audio(n, seed=arg, fs=int(fs))
    if kind == "noise":
        return (np.random.default_rng(arg).standard_normal(n) * 0.05).astype(np.float32)
    if kind == "silence":
        return np.zeros(n, np.float32)
    if kind == "dc":
        return np.full(n, 0.25, np.float32)
    raise ValueError(kind)


CASES = []
for n in (256, 1024, 4096):
    for f in (82.4, 110.0, 220.0, 329.6, 441.0, 1000.0, 3000.0, 60.0):
        CASES.append(("tone_n%d_f%g" % (n, f), n, 44100.0, "tone", f))
    for seed in range(5):
        CASES.append(("synth_n%d_s%d" % (n, seed), n, 44100.0, "synth", seed))
    CASES.append(("noise_n%d" % n, n, 44100.0, "noise", 7))
    CASES.append(("silence_n%d" % n, n, 44100.0, "silence", 0))
    CASES.append(("dc_n%d" % n, n, 48000.0, "dc", 0))
