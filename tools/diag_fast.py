"""diagnostic: fast tile kernel vs general tile kernel vs per-hop path, where do they differ?"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from zen_b200 import hps
from zen_b200.synth import synth_audio
FS, HOP, BETA = 44100.0, int(os.environ.get("DIAG_HOP", "1024")), 2.5
flags = int(os.environ.get("DIAG_FLAGS", "2"))
n_hops, n_streams = int(os.environ.get("DIAG_HOPS", "200")), 3
audio = np.stack([synth_audio(n_hops * HOP, seed=900 + s) for s in range(n_streams)])
x = torch.from_numpy(audio).cuda()
def run(no_fast):
    if no_fast: os.environ["ZEN_B200_NO_FAST"] = "1"
    else: os.environ.pop("ZEN_B200_NO_FAST", None)
    b = hps.HPRBatch(FS, HOP, BETA, flags)
    outs = b.process(x); torch.cuda.synchronize()
    r = [o.cpu().numpy() if o is not None else None for o in outs]
    b.close(); return r
f1, f2, g = run(False), run(False), run(True)
for o in range(3):
    if f1[o] is None: continue
    print("output", o, "fast deterministic:", np.array_equal(f1[o], f2[o]), " fast==general:", np.array_equal(f1[o], g[o]))
    d = (f1[o] != g[o]).reshape(n_streams, n_hops, HOP)
    per_hop = d.sum(axis=2)
    for s in range(n_streams):
        bad = np.nonzero(per_hop[s])[0]
        print(" stream", s, "differing hops", len(bad), "first", bad[:12], "samples/hop", per_hop[s][bad[:12]])
        if len(bad):
            h0 = bad[0]
            a = f1[o][s].reshape(n_hops, HOP)[h0]; c = g[o][s].reshape(n_hops, HOP)[h0]
            print("   hop", h0, "max abs diff", np.abs(a - c).max(), "scale", np.abs(c).max(), "idx of first diffs", np.nonzero(a != c)[0][:8])
