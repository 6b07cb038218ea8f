"""HBM-roofline check of the PCM16 decode / encode kernels: 512 streams x 60 s at 44.1 kHz.
CUDA events around each call (the wrappers launch on the legacy default stream, which is torch's), median of 9 after
warm-up; outputs are allocated before the timed calls (two buffers alternate in torch's caching allocator)."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
import torch
from zen_b200 import hps
n_streams, n = 512, 2646000
pcm = torch.randint(-32768, 32767, (n_streams, n), dtype=torch.int16, device="cuda")


def timed(fn, reps=9):
    keep = [fn(), fn()]            # both buffers of the ping-pong exist before timing
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        keep[0] = fn()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
        keep.reverse()
    return float(np.median(ms)), keep[0]


res = {}
ms, x = timed(lambda: hps.pcm16_decode_mono(pcm, 1))
res["decode_mono"] = {"ms": round(ms, 3), "algorithmic_GBps": round(n_streams * n * 6 / ms / 1e6, 1), "bytes": "2 in + 4 out per frame"}
ms, _ = timed(lambda: hps.pcm16_encode_normalized(x))
res["encode_normalized"] = {"ms": round(ms, 3), "algorithmic_GBps": round(n_streams * n * 10 / ms / 1e6, 1), "bytes": "4 (peak pass) + 4 + 2 per sample"}
# the two passes of the encode on their own (C ABI, legacy default stream)
from zen_b200 import _lib
L = _lib.lib()
q = torch.empty((n_streams, n), dtype=torch.int16, device="cuda")
pk = torch.zeros(n_streams, dtype=torch.float32, device="cuda")
ms, _ = timed(lambda: L.zen_pcm16_peaks_async(x.data_ptr(), x.stride(0), n_streams, n, pk.data_ptr(), None))
res["peak_pass"] = {"ms": round(ms, 3), "algorithmic_GBps": round(n_streams * n * 4 / ms / 1e6, 1)}
ms, _ = timed(lambda: L.zen_pcm16_encode_with_peaks_async(x.data_ptr(), x.stride(0), n_streams, n, pk.data_ptr(), q.data_ptr(), q.stride(0), None))
res["encode_pass"] = {"ms": round(ms, 3), "algorithmic_GBps": round(n_streams * n * 6 / ms / 1e6, 1)}
res["audio_s_per_s_decode_plus_encode"] = round(n_streams * 60.0 / ((res["decode_mono"]["ms"] + res["encode_normalized"]["ms"]) / 1e3))
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    res["hbm_copy_peak_GBps"] = peak
    res["decode_frac"] = round(res["decode_mono"]["algorithmic_GBps"] / peak, 3)
    res["encode_frac"] = round(res["encode_normalized"]["algorithmic_GBps"] / peak, 3)
except Exception:  # noqa: BLE001
    pass
print(json.dumps(res)); json.dump(res, open("gpurun_out/pcm_bench.json", "w"), indent=1)
