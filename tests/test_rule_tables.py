"""Rule tables of the product (host code behind the C ABI) against the reference's own host code.

Golden: tests/golden/rule_tables.npz (ndim 2..10) and rule_tables_hi.npz (11..15), written by tests/golden/make_rule_golden.py from
quad::Rule<double>::Init (reference cuda/pagani/quad/GPUquad/Rule.cuh:370-510).  Bit-exact.
"""
import os

import numpy as np
import pytest

import gpuintegration_b200 as pg

GOLDEN = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "rule_tables.npz")))
# ndim 11..15 (up to PGN_MAX_NDIM), same generator script, the arrays compared below only
GOLDEN.update(np.load(os.path.join(os.path.dirname(__file__), "golden", "rule_tables_hi.npz")))


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64) if a.dtype == np.float64 else a


@pytest.mark.parametrize("ndim", range(2, 16))
def test_tables_bit_exact(ndim):
    t = pg.rule_tables(ndim)
    assert pg.feval(ndim) == int(GOLDEN[f"n{ndim}_feval"])
    for ours, ref in [("weights", "weights"), ("scale", "scale"), ("norm", "norm"),
                      ("set_of_point", "set_of_point"), ("generators", "generators")]:
        assert np.array_equal(bits(t[ours]), bits(GOLDEN[f"n{ndim}_{ref}"])), (ndim, ours)


@pytest.mark.parametrize("ndim", [2, 5, 8, 10, 12, 15])
def test_point_codes_describe_generators(ndim):
    t = pg.rule_tables(ndim)
    code = t["point_code"]
    gens = t["generators"]
    lam = sorted(set(np.abs(gens[gens != 0]).tolist()))
    assert len(lam) == 5
    # code c -> generator value, as documented in include/pagani_b200.h
    first = {v: None for v in lam}
    for d in range(ndim):
        c = ((code >> np.uint64(4 * d)) & np.uint64(0xF)).astype(int)
        g = gens[d]
        assert np.all((c == 0) == (g == 0))
        assert np.all(((c % 2 == 0) & (c > 0)) == (g < 0))
    sets = (code >> np.uint64(60)).astype(np.int32)
    assert np.array_equal(sets, t["set_of_point"])
    assert np.all(np.diff(sets) >= 0) and sets[0] == 0 and sets[-1] == 8


def test_feval_formula_matches_quad_h():
    # quad.h:68-75; SURVEY 8: 5->273, 6->453, 8->1105, 10->2605
    assert [pg.feval(n) for n in (5, 6, 8, 10)] == [273, 453, 1105, 2605]
    with pytest.raises(ValueError):
        pg.feval(16)


def test_weights_integrate_constants():
    # rule 0 weights times set sizes sum to 1: the degree-9 rule integrates constants exactly
    for ndim in (2, 5, 8, 10):
        t = pg.rule_tables(ndim)
        counts = np.bincount(t["set_of_point"], minlength=9)
        assert abs(np.dot(counts, t["weights"][:, 0]) - 1.0) < 1e-12
        # null rules annihilate constants
        for r in (1, 2, 3, 4):
            assert abs(np.dot(counts, t["weights"][:, r])) < 1e-12
