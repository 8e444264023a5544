"""ctypes access to the CPU oracle (oracle/libpagani_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libpagani_oracle.so")
GOLDEN_RULES = os.path.join(ROOT, "tests", "golden", "rule_tables.npz")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class orc_iter(C.Structure):
    _fields_ = (
        [(n, C.c_uint64) for n in ("n", "nactive")]
        + [(n, C.c_double) for n in ("iter_estimate", "iter_errorest", "finished_estimate", "finished_errorest",
                                     "cum_estimate", "cum_errorest")]
        + [("heuristic_ran", C.c_int32), ("budget_overflow", C.c_int32)]
        + [(n, C.c_uint64) for n in ("dig_left", "dig_length", "dig_estimate", "dig_errorest_raw", "dig_errorest",
                                     "dig_axis", "dig_active")]
    )


class orc_result(C.Structure):
    _fields_ = [("estimate", C.c_double), ("errorest", C.c_double), ("neval", C.c_uint64),
                ("nregions", C.c_uint64), ("iters", C.c_uint64), ("status", C.c_int32),
                ("niter_records", C.c_int32)]


class cuhre_result(C.Structure):
    _fields_ = [("estimate", C.c_double), ("errorest", C.c_double), ("neval", C.c_uint64),
                ("nregions", C.c_uint64), ("status", C.c_int32), ("threads", C.c_int32), ("seconds", C.c_double)]


_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: make -C oracle oracle")
        _lib = C.CDLL(LIB)
        _lib.orc_evaluate.restype = C.c_int
        _lib.orc_integrate.restype = C.c_int
        _lib.cuhre_restated.restype = C.c_int
        _lib.orc_refine.restype = None
    return _lib


class Tables:
    """Rule tables of one dimension, straight from the reference's own host code (golden npz)."""

    def __init__(self, ndim):
        z = np.load(GOLDEN_RULES)
        self.ndim = ndim
        self.feval = int(z[f"n{ndim}_feval"])
        self.gens = np.ascontiguousarray(z[f"n{ndim}_generators"])
        self.weights = np.ascontiguousarray(z[f"n{ndim}_weights"])
        self.scale = np.ascontiguousarray(z[f"n{ndim}_scale"])
        self.norm = np.ascontiguousarray(z[f"n{ndim}_norm"])
        self.sop = np.ascontiguousarray(z[f"n{ndim}_set_of_point"])
        self.G = np.ascontiguousarray(z[f"n{ndim}_G"])

    def args(self):
        return (self.ndim, self.feval, self.gens.ctypes.data_as(_dp), self.weights.ctypes.data_as(_dp),
                self.scale.ctypes.data_as(_dp), self.norm.ctypes.data_as(_dp), self.sop.ctypes.data_as(_ip),
                self.G.ctypes.data_as(_dp))


def evaluate(family, ndim, left, length, lows=None, highs=None):
    """Per-region estimate / error / split axis of a region list [ndim][n].

    family 1..6: Genz; 0: the constant 15.37; -1: product of the coordinates.
    """
    t = Tables(ndim)
    left = np.ascontiguousarray(left, dtype=np.float64)
    length = np.ascontiguousarray(length, dtype=np.float64)
    n = left.shape[1]
    lows = np.zeros(ndim) if lows is None else np.ascontiguousarray(lows, dtype=np.float64)
    highs = np.ones(ndim) if highs is None else np.ascontiguousarray(highs, dtype=np.float64)
    est, err = np.zeros(n), np.zeros(n)
    axis = np.zeros(n, dtype=np.int32)
    rc = lib().orc_evaluate(C.c_int(family), *t.args(), left.ctypes.data_as(_dp), length.ctypes.data_as(_dp),
                            C.c_uint64(n), lows.ctypes.data_as(_dp), highs.ctypes.data_as(_dp),
                            est.ctypes.data_as(_dp), err.ctypes.data_as(_dp), axis.ctypes.data_as(_ip))
    if rc != 0:
        raise RuntimeError(f"orc_evaluate rc={rc}")
    return est, err, axis


def refine(est, err_raw, parent_est, epsrel, relerr=True):
    n = len(est)
    out = np.zeros(n)
    act = np.zeros(n, dtype=np.uint8)
    est = np.ascontiguousarray(est, dtype=np.float64)
    err_raw = np.ascontiguousarray(err_raw, dtype=np.float64)
    parent_est = np.ascontiguousarray(parent_est, dtype=np.float64)
    lib().orc_refine(est.ctypes.data_as(_dp), err_raw.ctypes.data_as(_dp), parent_est.ctypes.data_as(_dp),
                     C.c_uint64(n), C.c_double(epsrel), C.c_int(int(relerr)), out.ctypes.data_as(_dp),
                     act.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out, act


def classify_threshold(errors, iter_errorest, iter_finished_errorest, total_finished_errorest, last_estimate, epsrel):
    """Heuristic_classifier::classify (heuristic_classifier.cuh:362-451) on a host error list."""
    err = np.ascontiguousarray(errors, dtype=np.float64)
    flags = np.zeros(len(err), dtype=np.uint8)
    out = np.zeros(8)
    lib().orc_classify_threshold(err.ctypes.data_as(_dp), C.c_uint64(len(err)), C.c_double(iter_errorest),
                                 C.c_double(iter_finished_errorest), C.c_double(total_finished_errorest),
                                 C.c_double(last_estimate), C.c_double(epsrel),
                                 flags.ctypes.data_as(C.POINTER(C.c_uint8)), out.ctypes.data_as(_dp))
    return {"pass_mem": bool(out[0]), "pass_errorest_budget": bool(out[1]), "num_active": int(out[2]),
            "finished_errorest": out[3], "max_budget_perc_to_cover": out[4], "max_active_perc": out[5],
            "probes": int(out[6]), "threshold": out[7], "flags": flags}


def integrate(family, ndim, epsrel, epsabs=1e-40, lows=None, highs=None, relerr=True, mem_bytes=16e9,
              max_iters=700, max_regions=2_000_000, max_records=800):
    t = Tables(ndim)
    lows = np.zeros(ndim) if lows is None else np.ascontiguousarray(lows, dtype=np.float64)
    highs = np.ones(ndim) if highs is None else np.ascontiguousarray(highs, dtype=np.float64)
    res = orc_result()
    recs = (orc_iter * max_records)()
    rc = lib().orc_integrate(C.c_int(family), *t.args(), C.c_double(epsrel), C.c_double(epsabs),
                             lows.ctypes.data_as(_dp), highs.ctypes.data_as(_dp), C.c_int(int(relerr)),
                             C.c_double(mem_bytes), C.c_int(max_iters), C.c_uint64(max_regions), C.byref(res),
                             recs, C.c_int(max_records))
    if rc != 0:
        raise RuntimeError(f"orc_integrate rc={rc}")
    iters = [{name: getattr(recs[i], name) for name, _ in orc_iter._fields_}
             for i in range(min(res.niter_records, max_records))]
    return {name: getattr(res, name) for name, _ in orc_result._fields_}, iters


def cuhre(family, ndim, epsrel, epsabs=1e-100, maxeval=10**11, threads=0, time_cap_s=0.0):
    """Cuhre (restated, not libcuba) on the host cores; threads=0 -> all."""
    t = Tables(ndim)
    res = cuhre_result()
    rc = lib().cuhre_restated(C.c_int(family), *t.args(), C.c_double(epsrel), C.c_double(epsabs),
                              C.c_uint64(maxeval), C.c_int(threads), C.c_double(time_cap_s), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"cuhre_restated rc={rc}")
    return {name: getattr(res, name) for name, _ in cuhre_result._fields_}
