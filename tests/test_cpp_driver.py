"""examples/box_driver.cpp: a C++ program on include/ugf.h alone (the language of the reference's host code) - builds
with g++ against libugf.so, fails loudly without a GPU, and on a B200 passes its own conservation / collision-rate
checks."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


def build():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    subprocess.check_call(["make", "-C", EX, "-s", "-B", "box_driver"])
    return os.path.join(EX, "box_driver")


def test_cpp_driver_builds_and_refuses_to_run_without_a_gpu():
    import torch
    exe = build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "4", "200", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_runs_on_the_gpu():
    exe = build()
    r = subprocess.run([exe, "12", "60000", "40"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "box_driver: ok" in r.stdout
