"""BASELINE.json configs[2]: offline iterative 2-pass HPR (hop 4096 then 256, beta 2.5, hard mask) on a synthetic
10-minute 44.1 kHz signal; configs[3]: --soft-mask --sse hop 512 on 60 s.  Times zen_b200; the reference GPU path
on the same box took 25.5 s for configs[2] (tests/golden/ref_same_box_timings.json)."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from zen_b200 import hps
from zen_b200.synth import synth_audio
res = {}
a = synth_audio(600 * 44100, seed=3)
off = hps.HPRIOffline(44100.0, 4096, 256, 2.5, 2.5)
off.process(a[:44100 * 5])  # warm-up (module load, allocations)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); out = off.process(a); ts.append(time.perf_counter() - t0)
res["cfg3_offline_600s_host_in_out_ms"] = 1e3 * min(ts)
ad = torch.from_numpy(a).cuda()
off.process_device(ad); torch.cuda.synchronize()
ts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); o = off.process_device(ad); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
res["cfg3_offline_600s_device_resident_ms"] = 1e3 * min(ts)
res["cfg3_audio_s_per_s_host"] = 600.0 / (res["cfg3_offline_600s_host_in_out_ms"] * 1e-3)
res["cfg3_residual_all_zero"] = bool(np.all(out[2] == 0))
# config 4: real-time SSE + soft, hop 512, 60 s, one stream: batched kernel over the whole stream
b = hps.HPRBatch(44100.0, 512, 2.5, hps.OUTPUT_PERCUSSIVE, sse=True, soft=True)
x = torch.from_numpy(synth_audio(2646000, seed=4)[: (2646000 // 512) * 512][None, :]).cuda()
b.process(x); torch.cuda.synchronize()
ts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); b.process(x); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
res["cfg4_sse_hop512_60s_one_stream_ms"] = 1e3 * min(ts)
res["cfg4_audio_s_per_s"] = 60.0 / min(ts)
print(json.dumps(res, indent=1))
