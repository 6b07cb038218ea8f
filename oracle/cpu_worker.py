"""TEST INFRASTRUCTURE — one CPU worker of the reference-arm / cpu_baseline
measurement: runs real-time HPR (hop by hop, percussive output, hard mask, the
reference's CPU dataflow) over one synthetic stream on ONE core and prints a
JSON line with its own compute wall time.

kind "reference": oracle/_ref/libzen_ref.so, i.e. the reference's own hps.cu CPU
path; its IPP primitives are served by oracle/ref/ippstub (Intel IPP is not
installable here).  kind "port": oracle/libzen_oracle.so, the plain-C restatement.

    python oracle/cpu_worker.py <kind> <hop> <beta> <n_hops> <seed>
"""
import json
import os
import sys
import time

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from zen_b200.synth import FS, synth_audio  # noqa: E402


def main():
    kind, hop, beta, n_hops, seed = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    audio = synth_audio(n_hops * hop, seed=seed)
    if kind == "reference":
        from oracle import refbind as rb
        h = rb.RefHPR(rb.CPU, float(FS), hop, beta, rb.OUT_P, rb.CAUSAL, True)
    else:
        from oracle import oraclebind as ob
        h = ob.OracleHPR(ob.GEOM_CPU, float(FS), hop, beta, ob.OUT_P, ob.CAUSAL, True)
    h.run(audio[: 4 * hop], 4, want=(False, True, False))  # touch code and tables
    h.reset_buffers()
    t0 = time.perf_counter()
    out = h.run(audio, n_hops, want=(False, True, False))
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "audio_s": n_hops * hop / FS, "n_hops": n_hops,
                      "checksum": float(np.abs(out[1]).sum())}))


if __name__ == "__main__":
    main()
