"""Phase breakdown (device globaltimer) of one hop served by the resident real-time kernel."""
import ctypes, sys, time
sys.path.insert(0, ".")
import numpy as np
from zen_b200 import _lib, hps
from zen_b200.synth import synth_audio
hop = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
h = hps.HPR(44100.0, hop, 2.5, hps.OUTPUT_PERCUSSIVE, 0, True)
io = hps.IOGPU(hop)
audio = synth_audio(400 * hop, seed=1)
h.realtime_begin()
acc = []
for i in range(400):
    io.host_in[:] = audio[i * hop:(i + 1) * hop]
    t0 = time.perf_counter()
    h.process_hop_io(io.device_in, None, io.device_out, None)
    t1 = time.perf_counter()
    st = (ctypes.c_ulonglong * 16)()
    _lib.lib().zen_hpr_realtime_stamps(h._h, st)
    if i >= 100:
        s = list(st)
        acc.append([(t1 - t0) * 1e6] + [(s[k + 1] - s[k]) / 1e3 for k in range(8)] + [(s[0] - s[9]) / 1e3, (s[8] - s[9]) / 1e3, (s[11] - s[10]) / max(1.0, float(s[8] - s[9]))])
a = np.median(np.array(acc), axis=0)
names = ["host call us", "A load+window", "B fft fwd", "C split+mag", "F' H row", "E' decide", "G build", "G ifft", "G ola+emit", "pre", "kernel total", "SM clock GHz"]
for n, v in zip(names, a):
    print("%-14s %7.2f us" % (n, v))
h.close()
