"""Heuristic_classifier::classify on the engine through the stage-level C-ABI entry pgn_heuristic_classify:
the reference's own classifier known-answer test (test/kokkos/pagani/test_heuristic_classifier.cpp:32-103) and a
constructed list that takes the "exhausted right after a successful probe" path (the adopted flags are those of
the last probe, heuristic_classifier.cuh:404-437 / Workspace.cuh:104-106).  The CPU oracle runs the same two
cases in tests/test_oracle_golden.py; here engine and oracle must agree, device-side and host-side search.

(Passing on the GPU since the driver's round-1 run; part of every round-2 session's suite.)"""
import os

import numpy as np
import pytest

import oracle_cpu

pytestmark = pytest.mark.gpu


def run_case(pg, errors, last_estimate, epsrel, total_finished, host_search):
    n = len(errors)
    side = int(round(np.sqrt(n)))
    old = os.environ.get("PGN_HOST_SEARCH")
    os.environ["PGN_HOST_SEARCH"] = "1" if host_search else "0"
    try:
        with pg.Workspace(2) as ws:
            if side * side == n:
                ws.uniform_split(side)
            else:
                ws.import_regions(np.tile(np.arange(n) / n, (2, 1)), np.full((2, n), 1.0 / n))
            ws.set_estimates(np.full(n, last_estimate / n), errors, np.zeros(n, dtype=np.int32))
            ws.set_active(np.ones(n, dtype=np.uint8))
            r = ws.heuristic_classify(epsrel, last_estimate, float(np.sum(errors)), 0.0, total_finished)
            _, _, _, act = ws.fetch_estimates()
    finally:
        if old is None:
            del os.environ["PGN_HOST_SEARCH"]
        else:
            os.environ["PGN_HOST_SEARCH"] = old
    return r, act


@pytest.mark.parametrize("host_search", [False, True])
def test_reference_classifier_kat(pg, host_search):
    errors = np.array([.075, .99, .079, 101.96, 101.33, 1.93, 101.99])
    o = oracle_cpu.classify_threshold(errors, errors.sum(), 0., 4.2, 7000., 1e-3)
    r, act = run_case(pg, errors, 7000., 1e-3, 4.2, host_search)
    assert r["success"] and r["num_active"] == o["num_active"] == 4
    assert np.array_equal(act, o["flags"])
    assert r["probes"] == o["probes"]
    assert r["finished_errorest"] == pytest.approx(o["finished_errorest"], rel=1e-9)


@pytest.mark.parametrize("host_search", [False, True])
def test_flags_of_the_last_probe_are_adopted(pg, host_search):
    n = 4096
    errors = np.full(n, 1e-12)
    errors[:8] = 1.0
    errors[8:2048] = 5e-9
    tail = errors[2048:].sum()
    target = 4 * 1.5 * tail
    o = oracle_cpu.classify_threshold(errors, errors.sum(), 0., 0., target / 1e-3, 1e-3)
    assert o["probes"] == 20 and o["num_active"] == 2048
    r, act = run_case(pg, errors, target / 1e-3, 1e-3, 0.0, host_search)
    assert r["success"] and r["probes"] == 20
    assert r["num_active"] == 2048 and np.array_equal(act, o["flags"])
    assert 1e-12 < r["threshold"] < 5e-9          # the last probed threshold, not the reset iter_err / n
