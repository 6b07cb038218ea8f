"""The reference's OWN gtest sources, compiled unchanged against the drop-in headers ("the
reference's tests recompile unchanged"), and run on the GPU.

oracle/Makefile (target reftests) compiles /root/reference/libzen/{hps_gpu_public,mfilt,hps,fftw,box}.test.cu where
they lie - no copy, no edit - with tests/cpp/refshim/gtest/gtest.h standing in for googletest and
tests/cpp/refshim/cpu_backend.h supplying the Backend::CPU classes those translation units also instantiate (on top of
the oracle; the product has no CPU path).  The binaries live in oracle/_ref/reftests (git-ignored, they travel to the
GPU box, where /root/reference does not exist).

Not compiled: libzen/hps_cpu_public.test.cu - it only exercises HPRRealtime<CPU> / HPRIOffline<CPU>, the IPP path
that north_star keeps out of the product."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "reftests")
NAMES = ["hps_gpu_public", "mfilt", "hps", "fftw", "box"]
# libzen/CMakeLists.txt:82 - "#zen_unittest(box 1) # box filter testing fails": the reference disables box.test itself.
# Two of its five cases cannot hold for a moving AVERAGE, on NPP / IPP or anywhere: the frequency case expects 20
# (box.test.cu:203-234) and the CPU case expects 8 where the three GPU cases of the same file expect - and get - 32
# (box.test.cu:457-487 against :124-201).  The other three must pass.
KNOWN_REFERENCE_FAILURES = {"box": {"BoxFilterSmallSquareUnitTestGPU.Frequency", "BoxFilterSmallSquareUnitTestCPU.CausalTime"}}

def test_reference_tests_are_built():
    if os.path.isdir("/root/reference/libzen"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "reftests"])
    for n in NAMES:
        assert os.path.exists(os.path.join(BIN, n + ".test")), "oracle/_ref/reftests/%s.test missing: run `make -C oracle reftests`" % n


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_reference_gtest_binary(name):
    exe = os.path.join(BIN, name + ".test")
    if not os.path.exists(exe):
        pytest.skip("not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    ran = re.search(r"\[==========\] (\d+) tests ran", out)
    assert ran and int(ran.group(1)) > 0, out[-2000:]
    failed = set(re.findall(r"^\[  FAILED  \] ([\w.]+)$", out, flags=re.M))
    allowed = KNOWN_REFERENCE_FAILURES.get(name, set())
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, "reftest_%s.log" % name), "w") as f:
            f.write(out[-20000:])
    assert not (failed - allowed), (failed, out[-3000:])
    if not allowed:
        assert r.returncode == 0, out[-3000:]
    passed = set(re.findall(r"^\[       OK \] ([\w.]+)$", out, flags=re.M))
    assert len(passed) >= int(ran.group(1)) - len(allowed)
