"""The reference's OWN unit tests, compiled unmodified from /root/reference against this project's headers with
only the include path changed (tests/cpp/build_reference_tests.sh: include/pagani_b200/compat first), run on the GPU.

  RegionSplitting.cu, RegionFiltering.cu, RegionSampling.cu, Uniform_sub_division.cu   stage-level seams
  Easy_Integrals.cu, BoxIntegrals.cu, NanIntegral.cu, pagani_loop_test.cu,
  accuracy_improves_with_epsrel.cu                                                     Workspace::integrate
  do_integration_from_c.cu + test_integration_from_c.c                                  the C wrapper

The binaries are built where the reference tree exists (the build container) and travel to the GPU box."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "cpp", "_build", "ref_tests")
CATCH = ["RegionSplitting", "RegionFiltering", "RegionSampling", "Uniform_sub_division", "Easy_Integrals", "BoxIntegrals",
         "NanIntegral", "pagani_loop_test", "accuracy_improves_with_epsrel"]


REFBIN = os.path.join(HERE, "cpp", "_build", "ref_tests_reference")


def _run(name, timeout=900, bindir=BIN):
    exe = os.path.join(bindir, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} was not built (tests/cpp/build_reference_tests.sh needs the reference tree)")
    return subprocess.run([exe], capture_output=True, text=True, timeout=timeout, cwd=bindir)


def _catch_summary(out):
    """('passed', assertions, cases) or ('failed', cases_failed, assertions_failed, [failed CHECK locations])"""
    m = re.search(r"All tests passed \((\d+) assertions? in (\d+) test cases?\)", out)
    if m:
        return ("passed", int(m.group(1)), int(m.group(2)))
    cases = re.search(r"test cases:\s*(\d+) \|(?:\s*(\d+) passed \|)?\s*(\d+) failed", out)
    asserts = re.search(r"assertions:\s*(\d+) \|(?:\s*(\d+) passed \|)?\s*(\d+) failed", out)
    where = sorted(set(re.findall(r"([\w.]+\.cu:\d+): FAILED", out)))
    return ("failed", int(cases.group(3)) if cases else -1, int(asserts.group(3)) if asserts else -1, where)


# A reference program whose OWN build fails is no defect of the swap; the one known case:
# Uniform_sub_division.cu:64 expects 243 x (1/3)^5, summed on the host, to equal 1.0 exactly (it is 0.9999999999999968)
KNOWN_REFERENCE_FAILURES = {"Uniform_sub_division": ("failed", 1, 1, ["Uniform_sub_division.cu:64"])}


@pytest.mark.parametrize("name", [c for c in CATCH if c != "pagani_loop_test"])
def test_reference_catch2_program_behaves_like_the_reference_build(name):
    """Same Catch2 outcome as the program built against the reference's own headers and run on this GPU: all of its
    assertions pass -- or, where the reference's own build fails, the same assertions fail."""
    mine = _run(name)
    got = _catch_summary(mine.stdout)
    tail = (mine.stdout + mine.stderr)[-1500:]
    ref_exe = os.path.join(REFBIN, name)
    if os.path.exists(ref_exe):
        ref = _run(name, bindir=REFBIN)
        want = _catch_summary(ref.stdout)
        assert got == want, (got, want, tail)
        assert mine.returncode == ref.returncode
    if name in KNOWN_REFERENCE_FAILURES:
        assert got == KNOWN_REFERENCE_FAILURES[name], (got, tail)
    else:
        assert got[0] == "passed" and mine.returncode == 0, (got, tail)


def test_reference_loop_program_prints_what_the_reference_build_prints():
    """pagani_loop_test.cu: 10 000 integrate calls on one Workspace at tightening tolerances, a plain main() that
    prints every 100th result (estimate, errorest, nregions, chi_sq, status)."""
    mine = _run("pagani_loop_test")
    assert mine.returncode == 0
    lines = mine.stdout.strip().splitlines()
    assert len(lines) == 100 and lines[-1].startswith("10000 ")
    for l in lines:
        est, err, nregions, _chi, status = l.split(" ", 1)[1].split(",")
        assert status == "0" and abs(float(est) - 1.0) <= 1e-3 and int(nregions) > 0
    if os.path.exists(os.path.join(REFBIN, "pagani_loop_test")):
        ref = _run("pagani_loop_test", bindir=REFBIN)
        assert ref.returncode == 0
        rlines = ref.stdout.strip().splitlines()
        assert len(rlines) == len(lines)
        for a, b in zip(lines, rlines):  # same call, estimate, region count, status; error estimates to print precision
            ia, (ea, ra, na, _ca, sa) = a.split(" ", 1)[0], a.split(" ", 1)[1].split(",")
            ib, (eb, rb, nb, _cb, sb) = b.split(" ", 1)[0], b.split(" ", 1)[1].split(",")
            assert (ia, ea, na, sa) == (ib, eb, nb, sb), (a, b)
            assert float(ra) == pytest.approx(float(rb), rel=1e-4), (a, b)


def test_stateful_functors_and_snapshot_resume():
    """tests/cpp/test_stateful_resume.cu: functors with managed-memory state (the reference's cudaArray,
    cudaDynamicArray, Interp1D) through Workspace::integrate, and take_snapshot / load_snapshot -> new Workspace ->
    resume equals the uninterrupted run."""
    p = _run("test_stateful_resume")
    assert p.returncode == 0 and "0 failure(s)" in p.stdout, (p.stdout + p.stderr)[-2000:]


def test_reference_c_linkage_program_passes():
    p = _run("test_integration_from_c")
    assert p.returncode == 0, (p.stdout + p.stderr)[-1500:]
