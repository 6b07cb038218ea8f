"""TEST INFRASTRUCTURE — libzen/hps.bench.cu on the UNMODIFIED reference: per-hop time of HPRRealtime at 48 kHz,
hop 32...4096, including the mapped-memory copies (the region zen/fakert.h times), GPU path and CPU dataflow
(IPP stand-in).  Writes a JSON that tools/hps_bench.py places next to zen_b200's own numbers.

    gpurun -- python oracle/ref/ref_hps_bench.py gpurun_out/ref_hps_bench.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as rb  # noqa: E402
from zen_b200.synth import synth_audio  # noqa: E402

out = {}
for hop in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    n_h = 1000 if hop <= 1024 else 400
    a = synth_audio(n_h * hop, seed=hop, fs=48000)
    _, us = rb.fakert_latency(rb.GPU, 48000.0, hop, 2.0, a, n_h, warm=True)
    row = {"reference_gpu_p50_us": round(float(np.median(us)), 2)}
    _, us = rb.fakert_latency(rb.CPU, 48000.0, hop, 2.0, a, min(n_h, 100), warm=False)
    row["reference_cpu_standin_p50_us"] = round(float(np.median(us)), 2)
    out["hop%d" % hop] = row
    print(hop, row, flush=True)
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_hps_bench.json"
os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
json.dump(out, open(path, "w"), indent=1)
