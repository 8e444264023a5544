"""include/pagani_b200/fastmath.cuh: cos_pinned / cos_pinned_n and pow2_pinned / pow2_pinned_n must return the CUDA
math library's cos(double) and pow(double, 2.) bit for bit (the Genz oscillatory functor is built on it, and region-trajectory parity with
the reference needs f(x) bit-identical).  The comparison runs on the GPU (tests/cpp/test_fastmath.cu,
built by __graft_entry__.build()) over random, near-k*pi/2, boundary, huge and special inputs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_fastmath")


@pytest.mark.gpu
def test_cos_pinned_is_bit_identical_to_library_cos():
    assert os.path.exists(EXE), "tests/cpp/_build/test_fastmath missing: run __graft_entry__.build()"
    p = subprocess.run([EXE, "20000000", "pow"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert any(l.startswith("cos variant") and l.endswith("identical") for l in lines), p.stdout
    assert "pow2_pinned: identical" in lines, p.stdout
    # exp_pinned is opt-in and only reported (see fastmath.cuh): not asserted here
