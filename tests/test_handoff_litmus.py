"""Litmus / stress test of the cluster hand-off of the resident real-time kernel (tests/cpp/handoff_litmus.cu): the hop
stays in the leader's shared memory, the command word is stored into the other CTAs' shared memory after a CTA barrier,
and they pull the hop through the cluster window.  The product path puts a cluster-scope fence on both sides of the command
word (ZEN_B200_RT_FENCED=0 drops it); this shows that no reader sees a stale word either way, and what the fence costs
per round."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "handoff_litmus.cu")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "handoff_litmus")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= os.path.getmtime(SRC):
        return
    subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", SRC, "-o", EXE])


def test_litmus_compiles():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("cluster", [8, 4, 2])
def test_handoff_never_delivers_a_stale_hop(cluster):
    build()
    rounds = 2000000 if cluster == 8 else 500000
    r = subprocess.run([EXE, str(rounds), str(cluster)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads(r.stdout)
    assert res["plain"]["stale_words"] == 0 and res["fenced"]["stale_words"] == 0
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "handoff_litmus_c%d.json" % cluster), "w") as f:
        json.dump(res, f)


@pytest.mark.gpu
def test_fenced_resident_session_is_bit_identical(monkeypatch):
    """with and without fence.acq_rel.cluster on both sides of every hand-off: the same samples"""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from zen_b200 import hps
    rng = np.random.default_rng(5)
    x = (0.3 * rng.standard_normal(64 * 1024)).astype(np.float32)

    from zen_b200 import _lib
    L = _lib.lib()

    def run():
        h = hps.HPR(44100.0, 1024, 2.5, hps.OUTPUT_PERCUSSIVE, 0, True)
        h.realtime_begin()
        io = hps.IOGPU(1024)
        got = np.empty_like(x)
        for i in range(64):
            io.host_in[:] = x[i * 1024:(i + 1) * 1024]
            h.process_next_hop(io.device_in)
            assert L.zen_hpr_copy_percussive(h._h, io.device_out) == 0
            got[i * 1024:(i + 1) * 1024] = io.host_out
        h.realtime_end()
        h.close()
        return got

    a = run()
    monkeypatch.setenv("ZEN_B200_RT_FENCED", "0")
    b = run()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
