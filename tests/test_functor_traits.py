"""Host-only checks of the evaluation kernel's per-functor protocol (include/pagani_b200/evaluate.cuh: has_batch_fast,
has_batch_nocheck, the EvalTuning traits) and of the launch-time range promise of the oscillatory family
(genz::Oscillatory::fast_in_box): tests/cpp/test_functor_traits.cu is compiled by nvcc and run on the CPU.  The
promise decides which kernel instantiation a launch takes (with or without the per-point |s| < 2^31 checks of the
pinned cosine, the reference's cos() at common/cuda/integrands.cuh:1303-1320), so a wrong answer would be a silent
parity break on unusual integration volumes."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")


def test_functor_traits_and_range_promise(tmp_path):
    exe = str(tmp_path / "test_functor_traits")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-w", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_functor_traits.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == "ok"
