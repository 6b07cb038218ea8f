import sys; sys.path.insert(0, ".")
import numpy as np
from zen_b200 import _lib
from zen_b200.synth import synth_audio
hop = int(sys.argv[1]); n_h = int(sys.argv[2]); fused = int(sys.argv[3])
a = synth_audio(n_h * hop, seed=1)
perc = np.zeros(n_h * hop, np.float32); us = np.zeros(n_h)
_lib.check(_lib.lib().zen_fakert_run(44100.0, hop, 2.5, 0, a.ctypes.data, n_h, 50, fused, perc.ctypes.data, us.ctypes.data), "fakert")
print("p50 us", np.median(us))
