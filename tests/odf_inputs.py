"""seeded inputs of the onset detection tests (shared by the golden generator and the tests); hops of 256 samples"""
import numpy as np

from zen_b200.synth import synth_audio


def make_input(n_hops, kind, arg):
    n = n_hops * 256
    if kind == "synth":
        return synth_audio(n, seed=arg)
    if kind == "clicks":      # a click every `arg` samples on a quiet tone: what a percussive output looks like
        x = (0.01 * np.sin(2 * np.pi * 220.0 * np.arange(n) / 44100.0)).astype(np.float32)
        x[::arg] += 0.8
        x[1::arg] -= 0.5
        return x
    if kind == "noise":
        return (np.random.default_rng(arg).standard_normal(n) * 0.1).astype(np.float32)
    if kind == "silence":
        return np.zeros(n, np.float32)
    if kind == "tone":
        return (0.5 * np.sin(2 * np.pi * arg * np.arange(n) / 44100.0)).astype(np.float32)
    raise ValueError(kind)


CASES = [("synth_s%d" % s, 200, "synth", s) for s in range(4)]
CASES += [("clicks_5000", 300, "clicks", 5000), ("clicks_11025", 300, "clicks", 11025), ("noise", 150, "noise", 3),
          ("silence", 20, "silence", 0), ("tone_440", 100, "tone", 440.0), ("one_hop", 1, "noise", 9), ("three_hops", 3, "synth", 11)]
