"""The C-ABI library: loads, exports every symbol include/pagani_b200.h declares, fails loudly
without a GPU, and its host-only helpers behave (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpuintegration_b200 as pg
from gpuintegration_b200 import _capi as capi
from conftest import HAS_GPU, ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pagani_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(pgn_[a-z0-9_]+)\s*\(", text))
    # function-pointer typedefs are not exports
    names -= set(re.findall(r"\(\*\s*(pgn_[a-z0-9_]+)\s*\)", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.LIB_PATH)
    declared = declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in pagani_b200.h but not exported"
    # and the Python binding knows every one of them
    assert set(declared) == set(capi.SIGNATURES), set(declared) ^ set(capi.SIGNATURES)


def test_struct_layouts_match_header_sizes():
    # pgn_result_t mirrors numint::integration_result field for field (common/integration_result.hh:11-21)
    assert C.sizeof(capi.pgn_result_t) == 64
    assert C.sizeof(capi.pgn_opts) == 32
    assert C.sizeof(capi.pgn_integrate_args) == 32


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_gpu_is_a_loud_error_not_a_fallback():
    with pytest.raises(pg.PaganiError) as e:
        pg.Workspace(8)
    assert e.value.code == capi.PGN_ERR_NO_DEVICE
    with pytest.raises(pg.PaganiError):
        pg.fp64_peak_tflops()


def test_invalid_dimension_rejected():
    lib = capi.load()
    ctx = C.c_void_p()
    assert lib.pgn_create(C.byref(ctx), 1, None) == capi.PGN_ERR_INVALID
    assert lib.pgn_create(C.byref(ctx), 16, None) == capi.PGN_ERR_INVALID
    assert not lib.pgn_genz_available(1, 9)
    assert lib.pgn_genz_available(1, 8) and lib.pgn_genz_available(5, 10) and lib.pgn_genz_available(6, 10)


def test_digest_formula():
    rng = np.random.default_rng(7)
    for dtype in (np.float64, np.int32, np.uint8):
        a = (rng.standard_normal(1000) * 100).astype(dtype)
        w = {8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize]
        bits = a.view(w).astype(np.uint64)
        idx = np.arange(a.size, dtype=np.uint64)
        with np.errstate(over="ignore"):
            want = int(np.sum(bits * (np.uint64(2) * idx + np.uint64(1)), dtype=np.uint64))
        assert pg.digest64(a) == want


def test_genz_true_values_match_reference_constants():
    # integrands.cuh:1322 (F_1_8D), Easy_Integrals.cu:78,94,109,125,140
    tv = lambda f, n: pg._capi.load().pgn_genz_true_value(f, n)
    assert tv(1, 8) == pytest.approx(3.439557952183252e-05, rel=1e-12)
    assert tv(1, 5) == pytest.approx(0.020242422119901896863, rel=1e-12)
    assert tv(1, 6) == pytest.approx(-0.0013062949651908022873, rel=1e-11)
    # Easy_Integrals.cu:78 hard-codes 1.286889807581113e13; the closed form it is checked against to
    # 1e-3 (compute_genz_integrals.cuh:243-256) is 1.2868879901e13
    assert tv(2, 6) == pytest.approx(1.286889807581113e13, rel=1e-5)
    assert tv(3, 8) == 2.2751965817917756076e-10
    assert tv(4, 5) == pytest.approx(1.79132603674879e-06, rel=1e-12)
    assert tv(5, 8) == pytest.approx(2.425217625641885e-06, rel=1e-12)
    assert tv(6, 6) == pytest.approx(1.5477367885091207413e8, rel=1e-12)


def test_flops_per_eval_table():
    # SURVEY.md 8d contract figures
    fl = lambda f, n: pg._capi.load().pgn_genz_flops_per_eval(f, n)
    assert fl(1, 8) == 107 and fl(5, 10) == 142 and fl(6, 10) == 142
    assert fl(2, 6) == 940 and fl(4, 5) == 718 and fl(4, 8) == 1123
