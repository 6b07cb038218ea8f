"""Distribution and pattern of the per-hop latencies of zen_fakert_run on a resident session (back-to-back hops)."""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
from zen_b200 import _lib
from zen_b200.synth import MIXED_WAV_SAMPLES, synth_audio
L = _lib.lib()
FS, HOP, BETA = 44100, 1024, 2.5
n_h = 4000
mixed = synth_audio(MIXED_WAV_SAMPLES, seed=1)
a = np.tile(mixed, (n_h * HOP) // mixed.size + 1)[: n_h * HOP].copy()
if os.environ.get("LAT_SIGNAL") == "noise":      # no 158-hop period in the input: are the bursts tied to the data?
    a = (0.3 * np.random.default_rng(5).standard_normal(n_h * HOP)).astype(np.float32)
elif os.environ.get("LAT_SIGNAL") == "long":
    a = synth_audio(n_h * HOP, seed=3)
out = {}
if len(sys.argv) > 1:          # pin the calling thread to one core: are the bursts the guest's scheduler?
    os.sched_setaffinity(0, {int(sys.argv[1])})
    out["pinned_to_cpu"] = int(sys.argv[1])
for fused in (2, 3):
    perc = np.zeros(n_h * HOP, dtype=np.float32)
    us = np.zeros(n_h, dtype=np.float64)
    _lib.check(L.zen_fakert_run(float(FS), HOP, BETA, 0, a.ctypes.data, n_h, 1000, fused, perc.ctypes.data, us.ctypes.data), "run")
    q = {p: float(np.percentile(us, p)) for p in (1, 10, 25, 50, 75, 90, 95, 99)}
    slow = us > q[50] + 1.5
    idx = np.flatnonzero(slow)
    gaps = np.diff(idx)
    out[str(fused)] = {"pct": q, "mean": float(us.mean()), "slow_frac": float(slow.mean()), "slow_mean": float(us[slow].mean()) if slow.any() else None,
                       "gap_hist": np.bincount(np.minimum(gaps, 20)).tolist() if gaps.size else [], "first_slow": idx[:20].tolist(),
                       "burst_lengths": np.diff(np.flatnonzero(np.diff(np.concatenate([[0], slow.astype(np.int8), [0]])))).tolist()[::2][:30]}
print(json.dumps(out))
