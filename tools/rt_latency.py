"""p50 / p99 of the fakert region (zen/fakert.h:221-247) served by the resident kernel, per hop size, for the cluster
sizes of the split hop (ZEN_B200_RT_CLUSTER) and with ZEN_B200_RT_PUSH=0 (the kernel pulls the hop itself)."""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
from zen_b200 import _lib
from zen_b200.synth import synth_audio
L = _lib.lib()
hops = [int(x) for x in sys.argv[1:]] or [1024, 256]
out = {}
for hop in hops:
    n_h = 3000
    a = synth_audio(n_h * hop, seed=hop)
    ref = None
    for cluster, push, fused in (("1", "1", 2), ("4", "1", 2), ("8", "1", 2), ("4", "0", 2), ("4", "1", 3)):
        os.environ["ZEN_B200_RT_CLUSTER"] = cluster
        os.environ["ZEN_B200_RT_PUSH"] = push
        perc = np.zeros(n_h * hop, np.float32)
        us = np.zeros(n_h, np.float64)
        rc = L.zen_fakert_run(44100.0, hop, 2.5, 0, a.ctypes.data, n_h, 1000, fused, perc.ctypes.data, us.ctypes.data)
        if ref is None:
            ref = perc.copy()
        key = "hop%d_cluster%s_%s%s" % (hop, cluster, "push" if push == "1" else "pull", "_two_call" if fused == 3 else "")
        out[key] = {"rc": rc, "p50_us": round(float(np.median(us)), 2), "p99_us": round(float(np.percentile(us, 99)), 2),
                    "min_us": round(float(us.min()), 2), "same_as_cluster1": bool(np.array_equal(perc, ref))}
        print(key, out[key], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rt_latency.json", "w"), indent=1)
