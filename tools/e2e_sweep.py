"""Host-buffer (e2e) throughput of zen_hpr_batch_process_host as a function of the pipeline chunk size."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from zen_b200 import _lib, hps
from bench import synth_batch_device, FS, HOP, BETA
ns, n = 1024, 2583 * HOP
x = synth_batch_device(torch, ns, n, torch.device("cuda", 0), seed0=1)
h_in, h_out = hps.PinnedArray(ns, n), hps.PinnedArray(ns, n)
_lib.check(_lib.lib().zen_copy_to_host(h_in.ptr, x.data_ptr(), h_in.nbytes), "copy")
for mb in (16, 32, 64, 128, 256, 512):
    os.environ["ZEN_B200_CHUNK_MB"] = str(mb)
    b = hps.HPRBatch(float(FS), HOP, BETA, hps.OUTPUT_PERCUSSIVE)
    b.process_host(h_in.array, [None, h_out.array, None])
    t0 = time.perf_counter()
    for _ in range(3):
        b.process_host(h_in.array, [None, h_out.array, None])
    dt = (time.perf_counter() - t0) / 3
    print("chunk %4d MB: %.1f ms  %.0f audio-s/s  %.1f GB/s each way" % (mb, dt * 1e3, ns * n / FS / dt, ns * n * 4 / dt / 1e9), flush=True)
    b.close()
