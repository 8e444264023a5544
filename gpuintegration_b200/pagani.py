"""Host-side mirror of the reference's PAGANI interface, on top of the C ABI.

Reference seams mirrored here (all paths relative to the reference tree):
  Workspace<T,ndim>::integrate            cuda/pagani/quad/GPUquad/Workspace.cuh:52-70
  numint::integration_result              common/integration_result.hh:11-21
  quad::Volume<T,ndim>                    common/cuda/Volume.cuh:13-52
  Sub_regions<T,ndim>                     cuda/pagani/quad/GPUquad/Sub_regions.cuh:38-258
  Cubature_rules::apply_cubature_integration_rules   PaganiUtils.cuh:291-343
  Sub_regions_filter::filter              Sub_region_filter.cuh:85-139
  Sub_region_splitter::split              Sub_region_splitter.cuh:55-87

Integrands are device functors; from Python the pre-instantiated Genz families are available
(`Genz(family, ndim)`), arbitrary functors go through the C++ header include/pagani_b200/Workspace.cuh.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi as capi

GENZ_NAMES = {
    1: "oscillatory",
    2: "product_peak",
    3: "corner_peak",
    4: "gaussian",
    5: "c0",
    6: "discontinuous",
}


@dataclass
class IntegrationResult:
    """numint::integration_result (neval is filled here; the reference leaves it 0)."""

    estimate: float = 0.0
    errorest: float = 0.0
    neval: int = 0
    nregions: int = 0
    nFinishedRegions: int = 0
    status: int = -1
    lastPhase: int = -1
    chi_sq: float = 0.0
    iters: int = 0


class Volume:
    """quad::Volume: integration box, unit cube by default."""

    def __init__(self, ndim, lows=None, highs=None):
        self.ndim = ndim
        self.lows = np.zeros(ndim) if lows is None else np.asarray(lows, dtype=np.float64).copy()
        self.highs = np.ones(ndim) if highs is None else np.asarray(highs, dtype=np.float64).copy()
        if self.lows.shape != (ndim,) or self.highs.shape != (ndim,):
            raise ValueError("Volume: lows/highs must have ndim entries")


class Genz:
    """One of the pre-instantiated Genz integrands (family 1..6) at a given dimension."""

    def __init__(self, family, ndim):
        lib = capi.load()
        if not lib.pgn_genz_available(family, ndim):
            raise ValueError(f"Genz family {family} at ndim={ndim} is not instantiated in this build")
        self.family, self.ndim = family, ndim
        fn, user = C.c_void_p(), C.c_void_p()
        capi.check(lib.pgn_genz_evaluator(family, ndim, C.byref(fn), C.byref(user)), "pgn_genz_evaluator")
        self.fn, self.user = fn, user

    @property
    def name(self):
        return f"F_{self.family}_{self.ndim}D"

    @property
    def true_value(self):
        return capi.load().pgn_genz_true_value(self.family, self.ndim)

    @property
    def flops_per_eval(self):
        return capi.load().pgn_genz_flops_per_eval(self.family, self.ndim)


def feval(ndim):
    """pagani::CuhreFuncEvalsPerRegion (quad.h:68-75)."""
    v = capi.load().pgn_feval(ndim)
    if v < 0:
        raise ValueError("ndim must be in [2, 15]")
    return v


def rule_tables(ndim):
    """Host copy of the rule tables (Rule.cuh): dict of numpy arrays."""
    lib = capi.load()
    fe = feval(ndim)
    gens = np.zeros((ndim, fe))
    w = np.zeros((9, 5))
    sc = np.zeros((9, 5))
    no = np.zeros((9, 5))
    sop = np.zeros(fe, dtype=np.int32)
    code = np.zeros(fe, dtype=np.uint64)
    capi.check(
        lib.pgn_rule_tables_host(
            ndim,
            gens.ctypes.data_as(capi._dp),
            w.ctypes.data_as(capi._dp),
            sc.ctypes.data_as(capi._dp),
            no.ctypes.data_as(capi._dp),
            sop.ctypes.data_as(C.POINTER(C.c_int32)),
            code.ctypes.data_as(C.POINTER(C.c_uint64)),
        ),
        "pgn_rule_tables_host",
    )
    return {"generators": gens, "weights": w, "scale": sc, "norm": no, "set_of_point": sop, "point_code": code}


def digest64(a):
    """Order-sensitive 64-bit digest of an array (same formula as the device digests)."""
    a = np.ascontiguousarray(a)
    return int(capi.load().pgn_digest64(a.ctypes.data_as(C.c_void_p), a.dtype.itemsize, a.size))


def plan_rebalance(counts):
    """Host-only planner of the multi-GPU repartition -> (new_counts, [(src,dst,src_off,dst_off,count)])."""
    lib = capi.load()
    n = len(counts)
    c = (C.c_uint64 * n)(*[int(x) for x in counts])
    nc = (C.c_uint64 * n)()
    cap = 4 * n + 4
    tr = (C.c_uint64 * (5 * cap))()
    nt = C.c_int()
    capi.check(lib.pgn_plan_rebalance(c, n, nc, tr, cap, C.byref(nt)), "pgn_plan_rebalance")
    return list(nc), [tuple(int(tr[5 * k + j]) for j in range(5)) for k in range(nt.value)]


class Workspace:
    """Workspace<double, ndim>: owns the rule tables and all device memory of the engine."""

    def __init__(self, ndim, device=-1, mem_mode="parity", max_iters=0, trace=False, trace_arrays=False,
                 native_mem_fraction=0.0):
        self._lib = capi.load()
        self.ndim = ndim
        self._ctx = C.c_void_p()
        self._keep = []  # ctypes callbacks must outlive the context
        o = capi.pgn_opts(
            device=device,
            mem_mode=capi.PGN_MEM_NATIVE if mem_mode == "native" else capi.PGN_MEM_PARITY,
            max_iters=max_iters,
            trace=int(trace or trace_arrays),
            trace_arrays=int(trace_arrays),
            native_mem_fraction=native_mem_fraction,
        )
        capi.check(self._lib.pgn_create(C.byref(self._ctx), ndim, C.byref(o)), "pgn_create")

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.pgn_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- the reference's entry points -----------------------------------------------------------
    def integrate(self, integrand, epsrel, epsabs, vol=None, relerr_classification=True, subregions=False):
        """Workspace::integrate(integrand, epsrel, epsabs, vol, relerr_classification).

        subregions=True mirrors the overload that consumes the caller's current region list
        (Workspace.cuh:148-358; `iters` is then counted before the convergence test).
        """
        self._check_integrand(integrand)
        vol = vol or Volume(self.ndim)
        capi.check(
            self._lib.pgn_set_volume(self._ctx, vol.lows.ctypes.data_as(capi._dp), vol.highs.ctypes.data_as(capi._dp)),
            "pgn_set_volume",
        )
        a = capi.pgn_integrate_args(
            epsrel=epsrel,
            epsabs=epsabs,
            relerr_classification=int(relerr_classification),
            use_current_regions=int(subregions),
            iters_mode=capi.PGN_ITERS_BEFORE_CHECK if subregions else capi.PGN_ITERS_AFTER_SPLIT,
        )
        r = capi.pgn_result_t()
        capi.check(self._lib.pgn_integrate(self._ctx, integrand.fn, integrand.user, C.byref(a), C.byref(r)), "pgn_integrate")
        return IntegrationResult(r.estimate, r.errorest, r.neval, r.nregions, r.nFinishedRegions, r.status,
                                 r.lastPhase, r.chi_sq, r.iters)

    def trace(self):
        n = self._lib.pgn_trace_count(self._ctx)
        out = []
        for i in range(max(n, 0)):
            st = capi.pgn_iter_stats()
            capi.check(self._lib.pgn_trace_get(self._ctx, i, C.byref(st)), "pgn_trace_get")
            out.append(st.as_dict())
        return out

    # ---- Sub_regions ------------------------------------------------------------------------------
    def uniform_split(self, partitions_per_axis):
        capi.check(self._lib.pgn_regions_init_uniform(self._ctx, partitions_per_axis), "pgn_regions_init_uniform")

    def import_regions(self, left, length):
        left = np.ascontiguousarray(left, dtype=np.float64)
        length = np.ascontiguousarray(length, dtype=np.float64)
        if left.shape != length.shape or left.ndim != 2 or left.shape[0] != self.ndim:
            raise ValueError("regions must be [ndim][n] arrays")
        capi.check(
            self._lib.pgn_regions_import(self._ctx, left.ctypes.data_as(capi._dp), length.ctypes.data_as(capi._dp), left.shape[1]),
            "pgn_regions_import",
        )

    @property
    def num_regions(self):
        n = C.c_size_t()
        capi.check(self._lib.pgn_regions_count(self._ctx, C.byref(n)), "pgn_regions_count")
        return n.value

    def export_regions(self):
        n = self.num_regions
        left = np.zeros((self.ndim, n))
        length = np.zeros((self.ndim, n))
        got = C.c_size_t()
        capi.check(
            self._lib.pgn_regions_export(self._ctx, left.ctypes.data_as(capi._dp), length.ctypes.data_as(capi._dp), n, C.byref(got)),
            "pgn_regions_export",
        )
        return left, length

    def set_volume(self, vol):
        capi.check(
            self._lib.pgn_set_volume(self._ctx, vol.lows.ctypes.data_as(capi._dp), vol.highs.ctypes.data_as(capi._dp)),
            "pgn_set_volume",
        )

    # ---- stage-level seams ------------------------------------------------------------------------
    def apply_cubature_integration_rules(self, integrand):
        """Evaluate the current region list; returns (sum of estimates, sum of raw errors)."""
        self._check_integrand(integrand)
        e, r = C.c_double(), C.c_double()
        capi.check(self._lib.pgn_evaluate(self._ctx, integrand.fn, integrand.user, C.byref(e), C.byref(r)), "pgn_evaluate")
        return e.value, r.value

    def classify(self, epsrel, relerr_classification=True):
        st = capi.pgn_iter_stats()
        capi.check(self._lib.pgn_classify(self._ctx, epsrel, int(relerr_classification), C.byref(st)), "pgn_classify")
        return st.as_dict()

    def set_active(self, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        capi.check(self._lib.pgn_set_active(self._ctx, f.ctypes.data_as(C.POINTER(C.c_uint8)), f.size), "pgn_set_active")

    def set_estimates(self, estimate, errorest, axis):
        e = np.ascontiguousarray(estimate, dtype=np.float64)
        r = np.ascontiguousarray(errorest, dtype=np.float64)
        a = np.ascontiguousarray(axis, dtype=np.int32)
        capi.check(
            self._lib.pgn_set_estimates(self._ctx, e.ctypes.data_as(capi._dp), r.ctypes.data_as(capi._dp),
                                        a.ctypes.data_as(C.POINTER(C.c_int32)), e.size),
            "pgn_set_estimates",
        )

    def heuristic_classify(self, epsrel, last_estimate, iter_errorest, iter_finished_errorest=0.0,
                           total_finished_errorest=0.0):
        """Heuristic_classifier::classify (heuristic_classifier.cuh:362-451) on the current error list, as the
        reference's classifier test drives it; needs classify() or set_active() first."""
        r = capi.pgn_heuristic_result()
        capi.check(self._lib.pgn_heuristic_classify(self._ctx, epsrel, last_estimate, iter_errorest,
                                                    iter_finished_errorest, total_finished_errorest, C.byref(r)),
                   "pgn_heuristic_classify")
        return {"success": bool(r.success), "must_terminate": bool(r.must_terminate), "probes": r.probes,
                "num_active": r.num_active, "threshold": r.threshold, "finished_errorest": r.finished_errorest}

    def filter(self):
        m = C.c_size_t()
        capi.check(self._lib.pgn_filter(self._ctx, C.byref(m)), "pgn_filter")
        return m.value

    def split(self):
        capi.check(self._lib.pgn_split(self._ctx), "pgn_split")

    def fetch_estimates(self):
        n = self.num_regions
        est, err = np.zeros(n), np.zeros(n)
        axis = np.zeros(n, dtype=np.int32)
        act = np.zeros(n, dtype=np.uint8)
        capi.check(
            self._lib.pgn_fetch_estimates(self._ctx, est.ctypes.data_as(capi._dp), err.ctypes.data_as(capi._dp),
                                          axis.ctypes.data_as(C.POINTER(C.c_int32)), act.ctypes.data_as(C.POINTER(C.c_uint8)), n),
            "pgn_fetch_estimates",
        )
        return est, err, axis, act

    def fetch_parents(self):
        n = C.c_size_t()
        capi.check(self._lib.pgn_fetch_parents(self._ctx, None, None, None, 0, C.byref(n)), "pgn_fetch_parents")
        est, err = np.zeros(n.value), np.zeros(n.value)
        axis = np.zeros(n.value, dtype=np.int32)
        capi.check(
            self._lib.pgn_fetch_parents(self._ctx, est.ctypes.data_as(capi._dp), err.ctypes.data_as(capi._dp),
                                        axis.ctypes.data_as(C.POINTER(C.c_int32)), n.value, C.byref(n)),
            "pgn_fetch_parents",
        )
        return est, err, axis

    def device_info(self):
        sm, fr, tot = C.c_int(), C.c_size_t(), C.c_size_t()
        capi.check(self._lib.pgn_device_info(self._ctx, C.byref(sm), C.byref(fr), C.byref(tot)), "pgn_device_info")
        return {"sm_count": sm.value, "free_bytes": fr.value, "total_bytes": tot.value}

    # ---- multi-GPU --------------------------------------------------------------------------------
    def comm_init_nccl(self, unique_id, nranks, rank):
        buf = (C.c_uint8 * capi.PGN_NCCL_ID_BYTES).from_buffer_copy(bytes(unique_id))
        capi.check(self._lib.pgn_comm_init(self._ctx, buf, nranks, rank), "pgn_comm_init")

    def comm_init_host(self, nranks, rank, allreduce, alltoallv):
        """Collectives supplied by the launcher (e.g. torch.distributed gloo); see parallel.py."""
        ar = capi.pgn_allreduce_fn(allreduce)
        a2a = capi.pgn_alltoallv_fn(alltoallv)
        self._keep += [ar, a2a]
        capi.check(self._lib.pgn_comm_set_host_collective(self._ctx, nranks, rank, ar, a2a, None), "pgn_comm_set_host_collective")

    @property
    def cuda_stream(self):
        """Raw cudaStream_t of the engine (wrap with torch.cuda.ExternalStream for event timing)."""
        s = C.c_void_p()
        capi.check(self._lib.pgn_get_stream(self._ctx, C.byref(s)), "pgn_get_stream")
        return s.value or 0

    @property
    def launch_count(self):
        n = C.c_uint64()
        capi.check(self._lib.pgn_launch_count(self._ctx, C.byref(n)), "pgn_launch_count")
        return n.value

    def set_eval_variant(self, variant):
        """0 default, 1 shift/mask point loop, 2 byte-permute point loop (bit-identical results)."""
        capi.check(self._lib.pgn_set_eval_variant(self._ctx, variant), "pgn_set_eval_variant")

    def set_rebalance(self, max_over_mean=1.10, min_regions=4096):
        capi.check(self._lib.pgn_set_rebalance(self._ctx, max_over_mean, min_regions), "pgn_set_rebalance")

    def _check_integrand(self, integrand):
        if getattr(integrand, "ndim", None) != self.ndim:
            raise ValueError("integrand dimension does not match the Workspace")


def nccl_unique_id():
    buf = (C.c_uint8 * capi.PGN_NCCL_ID_BYTES)()
    capi.check(capi.load().pgn_nccl_unique_id(buf), "pgn_nccl_unique_id")
    return bytes(buf)


def fp64_peak_tflops(device=-1, iters=4096):
    v = C.c_double()
    capi.check(capi.load().pgn_bench_fp64_peak(device, iters, C.byref(v)), "pgn_bench_fp64_peak")
    return v.value


def hbm_copy_gbs(device=-1, nbytes=1 << 30, iters=10):
    v = C.c_double()
    capi.check(capi.load().pgn_bench_hbm_copy(device, nbytes, iters, C.byref(v)), "pgn_bench_hbm_copy")
    return v.value
