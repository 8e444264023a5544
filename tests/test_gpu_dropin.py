"""The C++ drop-in header (include/pagani_b200/Workspace.cuh) compiled against user-written functors:
runs tests/cpp/_build/test_dropin (built by __graft_entry__.build(), sources tests/cpp/*.cu)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_dropin")


def test_dropin_program():
    assert os.path.exists(BIN), "tests/cpp/_build/test_dropin missing: run __graft_entry__.build()"
    p = subprocess.run([BIN], capture_output=True, text=True, timeout=900)
    lines = p.stdout.splitlines()
    passed = [l for l in lines if l.startswith("PASS")]
    failed = [l for l in lines if l.startswith("FAIL")]
    assert p.returncode == 0 and not failed, p.stdout + p.stderr
    assert len(passed) >= 14
