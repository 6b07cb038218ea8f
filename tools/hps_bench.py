"""libzen/hps.bench.cu equivalent: HPRRealtime per-hop time at 48 kHz for hop 32...4096, including the
mapped-memory copies (the region zen/fakert.h times), three call styles of zen_b200.  The numbers of the unmodified
reference on the same pool (oracle/ref/ref_hps_bench.py -> tests/golden/ref_hps_bench_same_box.json) are merged in."""
import ctypes, json, sys
sys.path.insert(0, ".")
import numpy as np
from zen_b200 import _lib
from zen_b200.synth import synth_audio
L = _lib.lib()
REF = json.load(open("tests/golden/ref_hps_bench_same_box.json"))
out = {}
for hop in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    n_h = 1000 if hop <= 1024 else 400
    a = synth_audio(n_h * hop, seed=hop, fs=48000)
    row = {}
    for name, fused in (("zen_two_call", 0), ("zen_fused", 1), ("zen_resident", 2)):
        perc = np.zeros(n_h * hop, np.float32)
        us = np.zeros(n_h, np.float64)
        rc = L.zen_fakert_run(48000.0, hop, 2.0, 0, a.ctypes.data, n_h, 200, fused, perc.ctypes.data, us.ctypes.data)
        row[name + "_p50_us"] = round(float(np.median(us)), 2) if rc == 0 else None
    row.update(REF.get("hop%d" % hop, {}))
    row["hop_duration_us"] = round(1e6 * hop / 48000.0, 1)
    out["hop%d" % hop] = row
    print(hop, row, flush=True)
json.dump(out, open("gpurun_out/hps_bench.json", "w"), indent=1)
