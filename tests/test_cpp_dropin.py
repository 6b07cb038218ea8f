"""The C++ headers under zen_b200/include reproduce the libzen API surface:
a caller written against the reference (call patterns of zen/fakert.h,
zen/offline.h, libzen/mfilt.test.cu, libzen/hps.test.cu) compiles unchanged
and, on a GPU, behaves as the reference's tests expect."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "dropin_test.cu")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_test")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    deps = [SRC] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "zen_b200", "include")) for f in fs]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I" + os.path.join(ROOT, "zen_b200", "include"), SRC, "-o", EXE,
                           "-L" + os.path.join(ROOT, "zen_b200", "lib"), "-lzen_b200",
                           "-Xlinker", "-rpath," + os.path.join(ROOT, "zen_b200", "lib")])


def test_dropin_headers_compile():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_dropin_behaviour_on_gpu():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "dropin_test OK" in r.stdout, r.stdout + r.stderr
