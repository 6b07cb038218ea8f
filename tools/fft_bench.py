"""libzen/fftw.bench.cu equivalent: in-place C2C FFT, sizes 256...32768, forward + backward round trip on
device-resident data, ours (zen_fft_c2c) next to cuFFT (through torch.fft, which the reference's wrapper also calls)."""
import json, sys
sys.path.insert(0, ".")
import torch
from zen_b200 import hps
out = {}
for n in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    f = hps.FFTC2CWrapperGPU(n)
    x = (torch.rand(n, device="cuda") + 1j * torch.rand(n, device="cuda")).to(torch.complex64)
    f.fft_vec.copy_(x)
    def ours():
        f.forward(); f.backward()
    def cufft():
        torch.fft.ifft(torch.fft.fft(x), norm="forward")
    res = {}
    for name, fn in (("zen_b200_us", ours), ("cufft_torch_us", cufft)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) * 10.0, 2)   # us per forward+backward pair
        f.fft_vec.copy_(x)
    out["n%d" % n] = res
print(json.dumps(out, indent=1))
