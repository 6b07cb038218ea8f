"""Distribution and pattern of the per-hop latencies of zen_fakert_run on a resident session (back-to-back hops)."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from zen_b200 import _lib
from zen_b200.synth import MIXED_WAV_SAMPLES, synth_audio
L = _lib.lib()
FS, HOP, BETA = 44100, 1024, 2.5
n_h = 4000
mixed = synth_audio(MIXED_WAV_SAMPLES, seed=1)
a = np.tile(mixed, (n_h * HOP) // mixed.size + 1)[: n_h * HOP].copy()
out = {}
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
                       "head": [round(float(v), 2) for v in us[:40]]}
print(json.dumps(out))
