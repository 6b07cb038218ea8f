"""mfilt.bench.cu-style timing of the standalone MedianFilterGPU (N x N, L = 11, device-resident data) against the
NPP numbers of the unmodified reference on the same pool (tests/golden/ref_same_box_timings.json)."""
import json, sys
sys.path.insert(0, ".")
import torch
from zen_b200 import hps
ref = json.load(open("tests/golden/ref_same_box_timings.json"))["mfilt_bench_us"]
out = {}
for (T, F, L, d, cb) in [(1024, 1024, 11, 0, 0), (1024, 1024, 11, 0, 1), (1024, 1024, 11, 2, 0), (1024, 1024, 11, 2, 1),
                         (4096, 4096, 11, 0, 1), (4096, 4096, 11, 2, 1), (8192, 8192, 11, 0, 1), (8192, 8192, 11, 2, 1),
                         (16384, 16384, 11, 0, 1), (16384, 16384, 11, 2, 1),
                         (6, 4096, 46, 2, 1), (6, 4096, 3, 0, 1), (2, 16384, 186, 2, 1), (22, 1024, 12, 2, 1), (22, 1024, 11, 1, 1)]:
    src = torch.rand((T, F), device="cuda")
    dst = torch.zeros_like(src)
    f = hps.MedianFilterGPU(T, F, L, d, bool(cb))
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
    key = ("N%d_d%d_cb%d_us" % (T, d, cb)) if T == F else ("T%d_F%d_L%d_d%d_cb%d_us" % (T, F, L, d, cb))
    gbs = 8.0 * T * F / (us * 1e-6) / 1e9
    out[key] = {"zen_b200_us": round(us, 2), "npp_reference_us": ref.get(key), "algorithmic_GBps": round(gbs, 1)}
print(json.dumps(out, indent=1))
