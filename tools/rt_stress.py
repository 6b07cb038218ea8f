"""Soak test of the resident session's tagged transfers: many runs of the fakert loop through the resident kernel
(fused call and the reference's two calls) against the per-launch path on the same audio; every sample must be
identical.  Prints the number of differing samples per run and where they sit inside their 3-sample group."""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
from zen_b200 import _lib
from zen_b200.synth import synth_audio
L = _lib.lib()
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
out = {}
for hop in (1024, 256):
    n_h = 3000
    a = synth_audio(n_h * hop, seed=hop)
    ref = np.zeros(n_h * hop, np.float32)
    us = np.zeros(n_h, np.float64)
    assert L.zen_fakert_run(44100.0, hop, 2.5, 0, a.ctypes.data, n_h, 50, 1, ref.ctypes.data, us.ctypes.data) == 0
    bad_runs, bad_samples, lanes, p50 = 0, 0, [0, 0, 0], []
    for r in range(runs):
        got = np.zeros(n_h * hop, np.float32)
        fused = 3 if r % 4 == 3 else 2
        assert L.zen_fakert_run(44100.0, hop, 2.5, 0, a.ctypes.data, n_h, 50, fused, got.ctypes.data, us.ctypes.data) == 0
        p50.append(float(np.median(us)))
        d = np.flatnonzero(got.view(np.uint32) != ref.view(np.uint32))
        if d.size:
            bad_runs += 1
            bad_samples += int(d.size)
            for j in range(3):
                lanes[j] += int(np.count_nonzero((d % hop) % 3 == j))
            print("hop", hop, "run", r, "differs in", d.size, "samples, first at hop", int(d[0]) // hop, "offset", int(d[0]) % hop, flush=True)
    out["hop%d" % hop] = {"runs": runs, "hops_per_run": n_h, "runs_with_differences": bad_runs, "differing_samples": bad_samples,
                          "position_in_group": lanes, "p50_us_median_over_runs": round(float(np.median(p50)), 2)}
    print(out["hop%d" % hop], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rt_stress.json", "w"), indent=1)
