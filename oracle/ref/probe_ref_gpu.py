"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference (oracle/_ref/libzen_ref.so)
on a GPU box and writes golden vectors + same-box baseline timings.

    gpurun -- python oracle/ref/probe_ref_gpu.py gpurun_out/ref_probe

Outputs (copied into tests/golden/ by oracle/ref/collect_golden.py):
    median_gpu.npz   NPP median filter outputs (sentinel-prefilled dst)
    box_gpu.npz      NPP box filter outputs incl. inf inputs
    fft_gpu.npz      cuFFT outputs
    hpr_*.npz        separated audio + stage dumps of the reference GPU/CPU paths
    timings.json     same-box reference latencies / throughputs, box facts
"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as rb  # noqa: E402
from zen_b200.synth import synth_audio  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_probe"
os.makedirs(out_dir, exist_ok=True)
SENT = np.float32(-777.0)
T0 = time.time()
DRY = bool(os.environ.get("PROBE_DRY"))  # dry run on a GPU-less box: CPU backend, tiny sizes
if DRY:
    rb.GPU = rb.CPU


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %r" % (e,)


timings = {"box": {
    "nproc": os.cpu_count(),
    "mem": sh("free -g | head -2"),
    "cpu": sh("lscpu | egrep 'Model name|Socket|Thread|Core' "),
    "gpu": sh("nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv"),
    "ref_device_count": rb.lib().ref_device_count(),
}}
print(json.dumps(timings["box"], indent=1), flush=True)

# ---------------------------------------------------------------- median ----
med = {}
cases = []
for (T, F, Ls) in [((12), 32, (3, 6)), (6, 4096, (3, 46)), (22, 1024, (11, 12)), (2, 16384, (1, 186)),
                   (12, 2048, (6, 23)), (2, 8192, (1, 93)), (9, 9, (3,)), (10, 20, (5,)), (64, 64, (21,))]:
    for L in Ls:
        for d in (0, 1, 2):
            for cb in (0, 1):
                cases.append((T, F, L, d, cb))
for idx, (T, F, L, d, cb) in enumerate(cases):
    rng = np.random.default_rng(1000 + idx)
    src = rng.standard_normal((T, F)).astype(np.float32)
    if idx % 3 == 0:  # magnitude-like data with exact ties
        src = np.abs(src)
        src[rng.random((T, F)) < 0.1] = np.float32(0.25)
    key = "T%d_F%d_L%d_d%d_cb%d" % (T, F, L, d, cb)
    try:
        dst = rb.median_filter(rb.GPU, src, L, d, cb, dst_init=np.full((T, F), SENT, np.float32))
        med[key + "_seed"] = np.int64(1000 + idx)
        med[key + "_src"] = src
        med[key + "_dst"] = dst
    except ValueError:
        med[key + "_zgexception"] = np.int64(1)
np.savez_compressed(os.path.join(out_dir, "median_gpu.npz"), **med)
print("median cases:", len(cases), "t=%.1f" % (time.time() - T0), flush=True)

# ------------------------------------------------------------------- box ----
box = {}
bidx = 0
for (T, F, Ls) in [(12, 32, (3, 6)), (6, 512, (3, 46)), (22, 256, (11, 12)), (12, 512, (6, 23))]:
    for L in Ls:
        for d in (0, 1, 2):
            for variant in ("rand", "recip", "inf"):
                rng = np.random.default_rng(2000 + bidx)
                bidx += 1
                src = np.abs(rng.standard_normal((T, F))).astype(np.float32)
                if variant == "recip":
                    src = (np.float32(1.0) / (src * src)).astype(np.float32)
                if variant == "inf":
                    src = (np.float32(1.0) / (src * src)).astype(np.float32)
                    src[rng.random((T, F)) < 0.02] = np.inf
                key = "T%d_F%d_L%d_d%d_%s" % (T, F, L, d, variant)
                try:
                    dst = rb.box_filter(rb.GPU, src, L, d, dst_init=np.full((T, F), SENT, np.float32))
                    box[key + "_src"] = src
                    box[key + "_dst"] = dst
                except ValueError:
                    box[key + "_zgexception"] = np.int64(1)
np.savez_compressed(os.path.join(out_dir, "box_gpu.npz"), **box)
print("box cases:", bidx, "t=%.1f" % (time.time() - T0), flush=True)

# ------------------------------------------------------------------- fft ----
fftd = {}
for n in (64, 1024, 4096, 16384):
    rng = np.random.default_rng(3000 + n)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    fftd["n%d_x" % n] = x
    fftd["n%d_fwd" % n] = rb.fft(rb.GPU, x, inverse=False)
    fftd["n%d_inv" % n] = rb.fft(rb.GPU, x, inverse=True)
np.savez_compressed(os.path.join(out_dir, "fft_gpu.npz"), **fftd)

# ------------------------------------------------------------------- HPR ----
hpr_cases = [
    # name, backend, fs, hop, beta, flags, causality, copybord, sse, soft, n_hops, seed
    ("rt1024_cb", 0, 44100.0, 1024, 2.5, 7, 0, 1, 0, 0, 32, 11),
    ("rt1024_nocb", 0, 44100.0, 1024, 2.5, 7, 0, 0, 0, 0, 32, 11),
    ("rt1024_cb_ponly", 0, 44100.0, 1024, 2.5, 2, 0, 1, 0, 0, 32, 11),
    ("rt256_48k_cb", 0, 48000.0, 256, 2.0, 7, 0, 1, 0, 0, 100, 12),
    ("rt256_48k_nocb", 0, 48000.0, 256, 2.0, 7, 0, 0, 0, 0, 100, 12),
    ("ac256_48k_cb", 0, 48000.0, 256, 2.0, 7, 1, 1, 0, 0, 100, 12),
    ("ac256_48k_nocb", 0, 48000.0, 256, 2.0, 7, 1, 0, 0, 0, 100, 12),
    ("rt512_sse", 0, 44100.0, 512, 2.5, 7, 0, 1, 1, 1, 60, 13),
    ("rt512_soft", 0, 44100.0, 512, 2.5, 7, 0, 1, 0, 1, 60, 13),
    ("rt512_nocb", 0, 44100.0, 512, 2.5, 7, 0, 0, 0, 0, 60, 13),
    ("rt4096_cb", 0, 44100.0, 4096, 2.5, 7, 0, 1, 0, 0, 12, 14),
    ("rt2048_nocb", 0, 44100.0, 2048, 2.5, 7, 0, 0, 0, 0, 12, 15),
    ("ac4096_cb", 0, 44100.0, 4096, 2.5, 7, 1, 1, 0, 0, 12, 14),
    ("ac4096_nocb", 0, 44100.0, 4096, 2.5, 7, 1, 0, 0, 0, 12, 14),
    ("cpu_rt1024", 1, 44100.0, 1024, 2.5, 7, 0, 1, 0, 0, 32, 11),
]
for (name, be, fs, hop, beta, flags, caus, cb, sse, soft, n_hops, seed) in hpr_cases:
    if DRY:
        be, n_hops = 1, min(n_hops, 8)
    audio = synth_audio(n_hops * hop, seed=seed, fs=int(fs))
    d = {"params": np.array([be, fs, hop, beta, flags, caus, cb, sse, soft, n_hops, seed], dtype=np.float64),
         "audio_sha": np.frombuffer(bytes.fromhex(sha(audio)), dtype=np.uint8)}
    runs = []
    for rep in range(2):
        h = rb.RefHPR(be, fs, hop, beta, flags, caus, cb)
        if sse:
            h.use_sse_filter()
        if soft:
            h.use_soft_mask()
        outs = h.run(audio, n_hops)
        runs.append(outs)
        if rep == 0:
            d["geom"] = np.array([h.nwin, h.nfft, h.l_harm, h.l_perc, h.lag, h.stft_width, h.cola], dtype=np.float64)
            for f in ("s_mag", "harmonic_matrix", "percussive_matrix", "harmonic_mask", "percussive_mask",
                      "residual_mask", "reciprocal", "sliding_stft"):
                d["final_" + f] = h.get(f)
        h.close()
    d["harmonic"], d["percussive"], d["residual"] = runs[0]
    d["deterministic"] = np.array([int(np.array_equal(a, b)) for a, b in zip(runs[0], runs[1])])
    np.savez_compressed(os.path.join(out_dir, "hpr_%s.npz" % name), **d)
    print("t=%.1f" % (time.time() - T0), "hpr", name, "deterministic", d["deterministic"], "peak", [float(np.abs(o).max()) for o in runs[0]], flush=True)

# offline (HPRIOffline) small: 10 big hops + 11 samples like hps_gpu_public.test.cu:60-80
n_off = 10 * 4096 + 11
audio = synth_audio(n_off, seed=21)
for name, be, nocb, sse, soft in [] if DRY else [("offline_gpu", 0, 0, 0, 0), ("offline_gpu_nocb", 0, 1, 0, 0), ("offline_cpu", 1, 0, 0, 0),
                                   ("offline_gpu_soft", 0, 0, 0, 1), ("offline_gpu_sse", 0, 0, 1, 0)]:
    outs, ms = rb.offline_process(be, 44100.0, 4096, 256, 2.5, 2.5, audio, nocopybord=bool(nocb), sse=bool(sse), soft=bool(soft))
    np.savez_compressed(os.path.join(out_dir, "hpr_%s.npz" % name), harmonic=outs[0], percussive=outs[1], residual=outs[2],
                        params=np.array([be, 44100.0, 4096, 256, 2.5, 2.5, nocb, sse, soft, n_off, 21], dtype=np.float64))
    print(name, "ms", ms, "peaks", [float(np.abs(o).max()) for o in outs], flush=True)

# ----------------------------------------------------------- timings --------
lat = {}
mixed = synth_audio(161571, seed=1)
for hop in (256, 512, 1024, 2048, 4096):
    n_h = 20 if DRY else 2000
    a = np.tile(mixed, (n_h * hop) // mixed.size + 1)[: n_h * hop]
    _, us = rb.fakert_latency(rb.GPU, 44100.0, hop, 2.5, a, n_h, warm=True)
    lat["gpu_hop%d" % hop] = {"mean_us": float(us.mean()), "p50_us": float(np.median(us)), "p99_us": float(np.percentile(us, 99)), "n": n_h}
    print("ref GPU fakert hop", hop, lat["gpu_hop%d" % hop], flush=True)
for hop, n_h in ((256, 30), (1024, 30), (4096, 6)) if DRY else ((256, 300), (1024, 300), (4096, 60)):
    a = np.tile(mixed, (n_h * hop) // mixed.size + 1)[: n_h * hop]
    _, us = rb.fakert_latency(rb.CPU, 44100.0, hop, 2.5, a, n_h, warm=False)
    lat["cpu_hop%d" % hop] = {"mean_us": float(us.mean()), "p50_us": float(np.median(us)), "n": n_h}
    print("ref CPU(stand-in) fakert hop", hop, lat["cpu_hop%d" % hop], flush=True)
# config 4: sse + soft, hop 512, 60 s
a = synth_audio(26460 if DRY else 2646000, seed=4)
n_h = a.size // 512
t0 = time.time()
_, us = rb.fakert_latency(rb.GPU, 44100.0, 512, 2.5, a, n_h, sse=True, soft=True, warm=True)
lat["gpu_cfg4_sse_hop512"] = {"total_s": float(us.sum() * 1e-6), "mean_us": float(us.mean()), "n": n_h, "wall_s": time.time() - t0}
print("cfg4", lat["gpu_cfg4_sse_hop512"], flush=True)
timings["fakert"] = lat

# config 3: offline 2-pass on 600 s (reference GPU) -- and mixed.wav-sized
if DRY:
    print(json.dumps(timings, indent=1))
    sys.exit(0)
outs, ms = rb.offline_process(rb.GPU, 44100.0, 4096, 256, 2.5, 2.5, mixed)
timings["offline_gpu_mixed_ms"] = ms
timings["offline_gpu_mixed_residual_all_zero"] = bool(np.all(outs[2] == 0))
print("offline GPU mixed.wav-size ms", ms, flush=True)
outs, ms = rb.offline_process(rb.CPU, 44100.0, 4096, 256, 2.5, 2.5, mixed)
timings["offline_cpu_standin_mixed_ms"] = ms
print("offline CPU(stand-in) mixed.wav-size ms", ms, flush=True)
long = synth_audio(600 * 44100, seed=3)
t0 = time.time()
outs, ms = rb.offline_process(rb.GPU, 44100.0, 4096, 256, 2.5, 2.5, long)
timings["offline_gpu_600s_ms"] = ms
print("offline GPU 600 s ms", ms, "wall", time.time() - t0, flush=True)
del outs, long

# mfilt.bench-style NxN L=11 on device-resident data (mfilt.bench.cu:222-262)
mb = {}
for N in (1024, 4096, 8192):
    for d in (0, 2):
        for cb in (0, 1):
            mb["N%d_d%d_cb%d_us" % (N, d, cb)] = rb.lib().ref_median_filter_time(N, N, 11, d, cb, 5)
timings["mfilt_bench_us"] = mb
print(mb, flush=True)
# HPR-shaped matrices
for (T, F, L, d) in [(6, 4096, 46, 2), (6, 4096, 3, 0), (2, 16384, 186, 2), (22, 1024, 12, 2), (22, 1024, 11, 1)]:
    mb["T%d_F%d_L%d_d%d_cb1_us" % (T, F, L, d)] = rb.lib().ref_median_filter_time(T, F, L, d, 1, 50)
timings["mfilt_bench_us"] = mb

with open(os.path.join(out_dir, "timings.json"), "w") as f:
    json.dump(timings, f, indent=1)
print(json.dumps(timings, indent=1))
