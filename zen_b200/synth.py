"""Synthetic audio used by the tests, the fixtures and bench.py.

The reference ships `samples/mixed.wav` only as a Git-LFS pointer
(/root/reference/samples/mixed.wav:1-3), so every config in BASELINE.json is
restated on synthetic mono float32 audio in [-1, 1] at fs = 44100:
tones (harmonic content) + decaying noise bursts (percussive content) + a noise
floor (residual content).  Deterministic in `seed`.
"""
import numpy as np

FS = 44100
MIXED_WAV_SAMPLES = 161571  # README.md:98-103 of the reference


def synth_audio(n_samples: int, seed: int = 1, fs: int = FS) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / fs
    f1 = rng.uniform(110.0, 880.0)
    x = np.zeros(n_samples, dtype=np.float64)
    for k in range(1, 5):
        x += 0.4 * np.sin(2.0 * np.pi * f1 * k * t) / k
    # exponentially decaying white-noise bursts (tau = 5 ms) every 0.25-0.5 s
    pos = 0.0
    tau = 0.005
    blen = int(8 * tau * fs)
    env = np.exp(-np.arange(blen) / (tau * fs))
    while True:
        pos += rng.uniform(0.25, 0.5)
        i0 = int(pos * fs)
        if i0 >= n_samples:
            break
        m = min(blen, n_samples - i0)
        x[i0:i0 + m] += 0.5 * env[:m] * rng.standard_normal(m)
    x += 0.01 * rng.standard_normal(n_samples)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def synth_pcm16_roundtrip(x: np.ndarray) -> np.ndarray:
    """Quantise as a PCM16 wav would (libnyquist convention x/32767,
    vendor/libnyquist/include/libnyquist/Common.h:296-302)."""
    q = np.clip(np.round(x.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)
    return (q.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
