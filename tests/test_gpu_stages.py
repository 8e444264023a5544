"""Stage-level seams, mirroring the reference's unit tests on the same inputs:
Uniform_sub_division.cu, RegionSampling.cu, RegionSplitting.cu, RegionFiltering.cu
(test/cuda/pagani/), through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ndim,parts", [(3, 4), (5, 3), (10, 2)])
def test_uniform_sub_division(pg, ndim, parts):
    # Uniform_sub_division.cu:23-95: volumes sum to 1, each = 1/N, bounds inside [0,1)
    with pg.Workspace(ndim) as ws:
        ws.uniform_split(parts)
        left, length = ws.export_regions()
    n = parts**ndim
    assert left.shape == (ndim, n)
    vol = np.prod(length, axis=0)
    assert vol.sum() == pytest.approx(1.0, rel=1e-12)
    assert np.allclose(vol, 1.0 / n, rtol=1e-12)
    assert left.min() >= 0.0 and left.max() < 1.0
    # create_uniform_split index order (Sub_regions.cuh:11-36): axis 0 varies fastest
    idx = np.arange(n)
    for d in range(ndim):
        assert np.array_equal(left[d], ((idx // parts**d) % parts) * (1.0 / parts))
    assert len({tuple(c) for c in left.T}) == n


@pytest.mark.parametrize("parts", [5, 9, 14])
def test_region_sampling_constant(pg, parts):
    # RegionSampling.cu:48-171: per-region estimate of a constant == value / nregions
    with pg.Workspace(2) as ws:
        ws.uniform_split(parts)
        total, _ = ws.apply_cubature_integration_rules(pg.Genz(0, 2))
        est, err, axis, act = ws.fetch_estimates()
    n = parts * parts
    assert est == pytest.approx(np.full(n, 15.37 / n), rel=1e-6)
    assert total == pytest.approx(15.37, rel=1e-10)
    assert np.all(act == 1)
    assert np.all(err < 1e-12)


def test_split_all_regions_along_one_axis(pg):
    # RegionSplitting.cu:41-203: child i at index i, sibling at n + i; split axis halved, rest intact
    ndim, parts = 2, 5
    with pg.Workspace(ndim) as ws:
        ws.uniform_split(parts)
        left0, len0 = ws.export_regions()
        n = left0.shape[1]
        for axis in (1, 0):
            ws.import_regions(left0, len0)
            ws.set_estimates(np.zeros(n), np.zeros(n), np.full(n, axis, dtype=np.int32))
            ws.split()
            left1, len1 = ws.export_regions()
            assert left1.shape == (ndim, 2 * n)
            other = 1 - axis
            assert np.array_equal(left1[other, :n], left0[other]) and np.array_equal(left1[other, n:], left0[other])
            assert np.array_equal(len1[other, :n], len0[other]) and np.array_equal(len1[other, n:], len0[other])
            assert np.array_equal(left1[axis, :n], left0[axis])
            assert np.array_equal(left1[axis, n:], left0[axis] + len0[axis] / 2)
            assert np.array_equal(len1[axis, :n], len0[axis] / 2) and np.array_equal(len1[axis, n:], len0[axis] / 2)


def _filter_case(pg, n_parts, inactive):
    ndim = 2
    with pg.Workspace(ndim) as ws:
        ws.uniform_split(n_parts)
        left0, len0 = ws.export_regions()
        n = left0.shape[1]
        est = np.arange(n, dtype=np.float64) + 0.5
        err = np.arange(n, dtype=np.float64) * 0.25
        axis = (np.arange(n) % ndim).astype(np.int32)
        ws.set_estimates(est, err, axis)
        flags = np.ones(n, dtype=np.uint8)
        flags[list(inactive)] = 0
        ws.set_active(flags)
        m = ws.filter()
        keep = np.flatnonzero(flags)
        assert m == keep.size
        if m == 0:
            assert ws.num_regions == 0
            return
        left1, len1 = ws.export_regions()
        pest, perr, paxis = ws.fetch_parents()
        # order-preserving compaction; est/err become the PARENT arrays; split axes carried
        assert np.array_equal(left1, left0[:, keep]) and np.array_equal(len1, len0[:, keep])
        assert np.array_equal(pest, est[keep]) and np.array_equal(perr, err[keep])
        assert np.array_equal(paxis, axis[keep])
        ws.split()
        left2, len2 = ws.export_regions()
        assert left2.shape[1] == 2 * m
        for j, i in enumerate(keep):
            a = axis[i]
            assert left2[a, j] == left0[a, i] and left2[a, m + j] == left0[a, i] + len0[a, i] / 2
            assert len2[a, j] == len0[a, i] / 2 == len2[a, m + j]
            assert left2[1 - a, j] == left0[1 - a, i] == left2[1 - a, m + j]


@pytest.mark.parametrize("inactive", [(), (99,), (0,), tuple(range(100)), (15,), (0, 99), tuple(range(1, 100, 2))])
def test_filter_hundred_regions(pg, inactive):
    # RegionFiltering.cu:42-494 (all / none / first / last / region 15 / first and last inactive)
    _filter_case(pg, 10, inactive)


@pytest.mark.parametrize("inactive", [(), (0,)])
def test_filter_one_region_list(pg, inactive):
    # RegionFiltering.cu:130-165 uses a one-region list; the smallest grid here is 1 x 1
    _filter_case(pg, 1, inactive)


def test_filter_across_tile_boundaries(pg):
    # compaction offsets come from per-tile counts (1024 regions per tile): exercise several tiles
    ndim, parts = 2, 60  # 3600 regions
    rng = np.random.default_rng(3)
    with pg.Workspace(ndim) as ws:
        ws.uniform_split(parts)
        left0, len0 = ws.export_regions()
        n = left0.shape[1]
        est = rng.standard_normal(n)
        err = rng.random(n)
        axis = rng.integers(0, ndim, n).astype(np.int32)
        flags = (rng.random(n) < 0.37).astype(np.uint8)
        ws.set_estimates(est, err, axis)
        ws.set_active(flags)
        m = ws.filter()
        keep = np.flatnonzero(flags)
        assert m == keep.size
        left1, len1 = ws.export_regions()
        pest, perr, paxis = ws.fetch_parents()
        assert np.array_equal(left1, left0[:, keep]) and np.array_equal(pest, est[keep])
        assert np.array_equal(perr, err[keep]) and np.array_equal(paxis, axis[keep])


def test_classify_two_level_refinement_matches_oracle(pg):
    """RefineError (Phases.cuh:219-268) through the C ABI vs the CPU restatement, bit for bit
    (no libm involved)."""
    import oracle_cpu

    ndim = 5
    g = pg.Genz(5, ndim)
    with pg.Workspace(ndim) as ws:
        ws.uniform_split(2)
        ws.apply_cubature_integration_rules(g)
        ws.classify(1e-3)
        ws.filter()
        ws.split()
        pest, _, _ = ws.fetch_parents()
        ws.apply_cubature_integration_rules(g)
        est, err_raw, _, _ = ws.fetch_estimates()
        st = ws.classify(1e-3)
        _, err_ref, _, act = ws.fetch_estimates()
    o_err, o_act = oracle_cpu.refine(est, err_raw, pest, 1e-3)
    assert np.array_equal(err_ref.view(np.uint64), o_err.view(np.uint64))
    assert np.array_equal(act, o_act)
    assert st["nactive"] == int(o_act.sum())
    assert st["iter_errorest"] == pytest.approx(o_err.sum(), rel=1e-13)
    assert st["iter_estimate"] == pytest.approx(est.sum(), rel=1e-13)
