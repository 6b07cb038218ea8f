"""HBM-roofline check of the PCM16 decode / encode kernels: 512 streams x 60 s at 44.1 kHz."""
import json, sys, time
sys.path.insert(0, ".")
import torch
from zen_b200 import hps
n_streams, n = 512, 2646000
pcm = torch.randint(-32768, 32767, (n_streams, n), dtype=torch.int16, device="cuda")
res = {}
for name, fn, nbytes in (("decode_mono", lambda: hps.pcm16_decode_mono(pcm, 1), n_streams * n * 6),):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); 
    for _ in range(5): out = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    res[name] = {"ms": round(dt * 1e3, 3), "algorithmic_GBps": round(nbytes / dt / 1e9, 1)}
x = hps.pcm16_decode_mono(pcm, 1)
hps.pcm16_encode_normalized(x); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): q, p = hps.pcm16_encode_normalized(x)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
res["encode_normalized"] = {"ms": round(dt * 1e3, 3), "algorithmic_GBps": round(n_streams * n * 10 / dt / 1e9, 1), "bytes": "4 (peak pass) + 4 + 2 per sample"}
res["audio_s_per_s_decode_plus_encode"] = round(n_streams * 60.0 / ((res["decode_mono"]["ms"] + res["encode_normalized"]["ms"]) / 1e3))
print(json.dumps(res)); json.dump(res, open("gpurun_out/pcm_bench.json", "w"), indent=1)
