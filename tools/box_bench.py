"""Standalone BoxFilterGPU (wrap-padded moving average, libzen/box.h:194-213) on device-resident N x N data:
algorithmic bytes (8 per element) over the measured time, both axes."""
import json, sys
sys.path.insert(0, ".")
import torch
from zen_b200 import hps
out = {}
for (T, F, L, d) in [(4096, 4096, 11, 0), (4096, 4096, 11, 2), (16384, 16384, 11, 0), (16384, 16384, 11, 2), (16384, 16384, 47, 2),
                     (16384, 16384, 7, 0), (6, 4096, 46, 2), (6, 4096, 3, 0)]:
    src = torch.rand((T, F), device="cuda")
    dst = torch.zeros_like(src)
    f = hps.BoxFilterGPU(T, F, L, d)
    for _ in range(3):
        f.filter(src, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        f.filter(src, dst)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    out["T%d_F%d_L%d_d%d" % (T, F, L, d)] = {"us": round(us, 2), "algorithmic_GBps": round(8.0 * T * F / (us * 1e-6) / 1e9, 1)}
print(json.dumps(out, indent=1))
