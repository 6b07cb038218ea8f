"""Throughput of the batched kernel over the variants of the path (hop, outputs, mask kind, causality, border):
audio-seconds per second on one GPU with the batch resident in HBM."""
import json, sys
sys.path.insert(0, ".")
import torch
from zen_b200 import hps
from bench import synth_batch_device, FS
cases = [
    # hop, flags, causal, nocopybord, sse, soft, streams, seconds
    (1024, 2, True, False, False, False, 1024, 30),
    (1024, 7, True, False, False, False, 1024, 30),
    (1024, 2, True, True, False, False, 1024, 30),
    (1024, 2, True, False, False, True, 1024, 30),
    (1024, 7, True, False, False, True, 1024, 30),
    (1024, 3, True, False, True, False, 1024, 30),
    (256, 2, True, False, False, False, 1024, 30),
    (256, 2, False, False, False, False, 1024, 30),
    (512, 2, True, False, False, False, 1024, 30),
    (2048, 2, True, False, False, False, 1024, 30),
    (4096, 7, False, False, False, False, 1024, 30),
    (4096, 2, True, False, False, True, 512, 30),
]
out = []
for (hop, flags, causal, nocb, sse, soft, ns, sec) in cases:
    n = (sec * FS // hop) * hop
    x = synth_batch_device(torch, ns, n, torch.device("cuda", 0), seed0=1)
    b = hps.HPRBatch(float(FS), hop, 2.5, flags, causal=causal, nocopybord=nocb, sse=sse, soft=soft)
    outs = [torch.empty_like(x) if flags & (1 << o) else None for o in range(3)]
    for _ in range(2):
        b.process(x, outs)
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        b.process(x, outs)
        ms.append(b.last_kernel_ms)
    r = {"hop": hop, "flags": flags, "causal": causal, "nocopybord": nocb, "sse": sse, "soft": soft, "streams": ns,
         "kernel_ms": round(min(ms), 2), "audio_s_per_s": round(ns * n / FS / (min(ms) * 1e-3)), "Mhops_per_s": round(ns * (n // hop) / min(ms) / 1e3, 2)}
    out.append(r)
    print(r, flush=True)
    b.close()
    del x, outs
json.dump(out, open("gpurun_out/batch_sweep.json", "w"), indent=1)
