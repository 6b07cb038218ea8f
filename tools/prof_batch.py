import sys; sys.path.insert(0, ".")
import torch, numpy as np
from zen_b200 import hps
from bench import synth_batch_device, fakert_hops, FS, HOP, BETA
n_streams, seconds = int(sys.argv[1]), int(sys.argv[2])
n = fakert_hops(seconds * FS, HOP) * HOP
x = synth_batch_device(torch, n_streams, n, torch.device("cuda", 0), seed0=1000)
out = torch.empty_like(x)
b = hps.HPRBatch(float(FS), HOP, BETA, hps.OUTPUT_PERCUSSIVE)
for _ in range(int(sys.argv[3])):
    b.process(x, [None, out, None])
torch.cuda.synchronize()
print("kernel ms", b.last_kernel_ms)
