"""Phase breakdown (device globaltimer) of one hop served by the resident real-time kernel."""
import ctypes, os, sys, time
os.environ["ZEN_B200_RT_STAMPS"] = "1"
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
acc1 = []
for i in range(400):
    io.host_in[:] = audio[i * hop:(i + 1) * hop]
    t0 = time.perf_counter()
    h.process_hop_io(io.device_in, None, io.device_out, None)
    t1 = time.perf_counter()
    st = (ctypes.c_ulonglong * 16)()
    _lib.lib().zen_hpr_realtime_stamps(h._h, st)
    st1 = (ctypes.c_ulonglong * 16)()
    _lib.lib().zen_hpr_realtime_stamps_rank1(h._h, st1)
    if i >= 100:
        s = list(st)
        ghz = (s[11] - s[10]) / max(1.0, float(s[12] - s[9]))  # SM cycles per globaltimer ns over the whole hop
        cyc = 1e3 * ghz                                        # cycles per us
        acc.append([(t1 - t0) * 1e6] + [(s[k + 1] - s[k]) / cyc for k in range(8)] + [(s[0] - s[10]) / cyc, (s[8] - s[10]) / cyc, ghz,
                   (s[13] - s[4]) / cyc, (s[14] - s[13]) / cyc, (s[5] - s[14]) / cyc])
        r = list(st1)
        if r[12] > r[9] > 0:
            g1 = (r[11] - r[10]) / max(1.0, float(r[12] - r[9]))
            c1 = 1e3 * g1
            acc1.append([(r[9] - s[9]) * 1e-3, (r[12] - s[12]) * 1e-3] + [(r[k + 1] - r[k]) / c1 for k in range(5)] + [(r[11] - r[5]) / c1, (r[0] - r[10]) / c1])
a = np.median(np.array(acc), axis=0)
names = ["host call us", "A load+window", "B fft fwd", "C split+mag", "F' H row", "E' decide", "G build", "G ifft", "G ola+emit", "pre", "kernel total", "SM clock GHz",
         "E'.1 H+thresholds (thread 0)", "E'.2 counting (thread 0)", "E'.3 rest + barrier"]
for n, v in zip(names, a):
    print("%-30s %7.2f us" % (n, v))
h.close()
if acc1:
    b = np.median(np.array(acc1), axis=0)
    print("-- CTA 1 of the cluster (globaltimer relative to the leader)")
    for n, v in zip(["saw the command after the leader detected the hop", "finished after the leader", "A", "B fft fwd (incl. pulling the hop from the leader)", "C split+mag", "F'", "E' decide", "masked spectrum to its owner + cluster barrier", "pre"], b):
        print("%-52s %7.2f us" % (n, v))
