"""TEST INFRASTRUCTURE — copies the subset of gpurun_out/ref_probe (written on
the B200 box by oracle/ref/probe_ref_gpu.py from the UNMODIFIED reference) that
the tests use into tests/golden/.

The reference GPU path is only run-to-run deterministic for stft_width == 2
(hop >= 2048): its ring shift is an overlapping thrust::copy
(libzen/hps.cu:469-470), a data race on the GPU.  Separated audio is therefore
kept only for the deterministic configs; for the others we keep the final-hop
stage dumps (s_mag -> harmonic/percussive matrices -> masks), which are
self-consistent snapshots and pin every per-stage computation.
"""
import hashlib
import json
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_probe"
dst = sys.argv[2] if len(sys.argv) > 2 else "tests/golden"
os.makedirs(dst, exist_ok=True)


def save(name, **arrs):
    np.savez_compressed(os.path.join(dst, name), **arrs)
    print(name, "%.0f KB" % (os.path.getsize(os.path.join(dst, name)) / 1024))


# median: keep src seeds + dst for all cases up to 64 K elements, hashes would lose the untouched-cell map
med = np.load(os.path.join(src, "median_gpu.npz"))
keep = {}
for k in med.files:
    if k.endswith("_src") or k.endswith("_dst"):
        a = med[k]
        if a.size > 16384:
            # big cases: the input is regenerated from the seed (tests/golden_inputs.py),
            # the NPP output is pinned by its SHA-256 (bit-exact comparison is all we need)
            if k.endswith("_dst"):
                keep[k[:-4] + "_dstsha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
            continue
    keep[k] = med[k]
save("npp_median.npz", **keep)
box = np.load(os.path.join(src, "box_gpu.npz"))
save("npp_box.npz", **{k: box[k] for k in box.files})
fft = np.load(os.path.join(src, "fft_gpu.npz"))
save("cufft.npz", **{k: fft[k] for k in fft.files if not k.startswith("n16384")})

audio_cases = ["rt4096_cb", "rt2048_nocb", "ac4096_cb", "ac4096_nocb"]
stage_cases = ["rt1024_cb", "rt1024_nocb", "rt256_48k_cb", "ac256_48k_nocb", "rt512_sse", "rt512_soft"]
for c in audio_cases:
    d = np.load(os.path.join(src, "hpr_%s.npz" % c))
    save("ref_gpu_audio_%s.npz" % c, params=d["params"], geom=d["geom"], audio_sha=d["audio_sha"],
         harmonic=d["harmonic"], percussive=d["percussive"], residual=d["residual"], deterministic=d["deterministic"])
for c in stage_cases:
    d = np.load(os.path.join(src, "hpr_%s.npz" % c))
    keep = {k: d[k] for k in d.files if k.startswith("final_") and k != "final_sliding_stft"}
    stft = d["final_sliding_stft"]
    lag = int(d["geom"][4])
    keep["final_stft_row"] = stft[stft.shape[0] - lag]   # the consumed row only
    save("ref_gpu_stages_%s.npz" % c, params=d["params"], geom=d["geom"], deterministic=d["deterministic"], **keep)
for c in ["offline_gpu", "offline_gpu_nocb", "offline_cpu"]:
    d = np.load(os.path.join(src, "hpr_%s.npz" % c))
    # pass 2 (hop 256, stft_width 22) of the GPU path is racy: keep harmonic (pass 1) + residual only
    arrs = dict(params=d["params"], harmonic=d["harmonic"], residual_all_zero=np.array(int(np.all(d["residual"] == 0))))
    if c == "offline_cpu":
        arrs["percussive"] = d["percussive"]
    save("ref_%s.npz" % c, **arrs)
t = json.load(open(os.path.join(src, "timings.json")))
json.dump(t, open(os.path.join(dst, "ref_same_box_timings.json"), "w"), indent=1)
