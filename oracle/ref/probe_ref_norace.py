"""TEST INFRASTRUCTURE - end-to-end audio goldens of the reference GPU path (thrust + cuFFT + NPP) at the headline
hops, from the reference built with the one-hunk race fix (oracle/Makefile: ref_norace, oracle/ref/norace.patch).

    ZEN_REF_SO=oracle/_ref/libzen_ref_norace.so python oracle/ref/probe_ref_norace.py tests/golden      (on a GPU box)

The unmodified reference is nondeterministic for stft_width > 2 (overlapping thrust::copy, libzen/hps.cu:469-470), so
round 1 could pin hop <= 1024 only stage by stage.  With the shift done through a temporary every run is repeated and
bit-compared; a case is written only if the two runs agree.  Besides the separated audio each golden carries the
consumed row of the reference's hard masks for every hop (bit-packed), so that the parity test can tell a threshold
flip (our FFT and cuFFT differ in the last bits) from an error."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("ZEN_REF_SO", os.path.join(ROOT, "oracle", "_ref", "libzen_ref_norace.so"))
from oracle import refbind as rb  # noqa: E402
from zen_b200.synth import synth_audio  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_norace"
os.makedirs(out_dir, exist_ok=True)
CASES = [
    # name, fs, hop, beta, flags, copy_bord, n_hops, seed, sse, soft      (HPRRealtime<GPU>: causal)
    ("rt1024_cb", 44100.0, 1024, 2.5, 7, 1, 40, 31, 0, 0),
    ("rt512_cb", 44100.0, 512, 2.5, 7, 1, 60, 32, 0, 0),
    ("rt256_cb", 44100.0, 256, 2.5, 7, 1, 100, 33, 0, 0),
    ("rt1024_nocb", 44100.0, 1024, 2.5, 7, 0, 40, 34, 0, 0),
    # the soft-mask and SSE variants (BASELINE.json configs[3] is hop 512 with both switches)
    ("rt512_sse_soft", 44100.0, 512, 2.5, 3, 1, 60, 35, 1, 1),
    ("rt1024_soft", 44100.0, 1024, 2.5, 7, 1, 40, 36, 0, 1),
    ("rt1024_sse", 44100.0, 1024, 2.5, 3, 1, 40, 37, 1, 0),
]
ONLY = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None


def one_run(fs, hop, beta, flags, cb, audio, n_hops, sse=0, soft=0):
    h = rb.RefHPR(rb.GPU, fs, hop, beta, flags, rb.CAUSAL, cb)
    if sse:
        h.use_sse_filter()
    if soft:
        h.use_soft_mask()
    row = h.stft_width - h.lag
    outs = [np.zeros(n_hops * hop, np.float32) for _ in range(3)]
    masks = {"harmonic_mask": [], "percussive_mask": []}
    for i in range(n_hops):
        h.process_next_hop(audio[i * hop:(i + 1) * hop])
        for o, nm in enumerate(("harmonic_out", "percussive_out", "residual_out")):
            outs[o][i * hop:(i + 1) * hop] = h.get(nm)[:hop]
        for nm in masks:
            masks[nm].append(np.packbits(h.get(nm)[row] != 0))
    geom = np.array([h.nwin, h.nfft, h.l_harm, h.l_perc, h.lag, h.stft_width, h.cola], dtype=np.float64)
    h.close()
    return outs, {k: np.stack(v) for k, v in masks.items()}, geom


for name, fs, hop, beta, flags, cb, n_hops, seed, sse, soft in CASES:
    if ONLY and name not in ONLY:
        continue
    audio = synth_audio(n_hops * hop, seed=seed, fs=int(fs))
    a = one_run(fs, hop, beta, flags, cb, audio, n_hops, sse, soft)
    b = one_run(fs, hop, beta, flags, cb, audio, n_hops, sse, soft)
    same = all(np.array_equal(x, y) for x, y in zip(a[0], b[0])) and all(np.array_equal(a[1][k], b[1][k]) for k in a[1])
    print(name, "deterministic", same, "peaks", [float(np.abs(o).max()) for o in a[0]], flush=True)
    if not same:
        continue
    np.savez_compressed(os.path.join(out_dir, "ref_norace_%s.npz" % name),
                        params=np.array([fs, hop, beta, flags, cb, n_hops, seed, sse, soft], dtype=np.float64), geom=a[2],
                        audio_sha=np.frombuffer(hashlib.sha256(audio.tobytes()).digest(), dtype=np.uint8),
                        harmonic=a[0][0], percussive=a[0][1], residual=a[0][2],
                        harmonic_mask_bits=a[1]["harmonic_mask"], percussive_mask_bits=a[1]["percussive_mask"])
