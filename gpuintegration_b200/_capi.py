"""ctypes binding of libpagani_b200.so (the C ABI declared in include/pagani_b200.h).

There is deliberately no fallback: if the shared library is missing this module raises, and if
there is no sm_100 GPU every entry point that needs one returns PGN_ERR_NO_DEVICE, which the
wrappers turn into an exception.
"""
import ctypes as C
import os

PGN_MAX_NDIM = 15
PGN_NSETS = 9
PGN_NRULES = 5
PGN_NCODES = 11
PGN_NCCL_ID_BYTES = 128

PGN_OK = 0
PGN_ERR_INVALID = -1
PGN_ERR_CUDA = -2
PGN_ERR_NOMEM = -3
PGN_ERR_NO_DEVICE = -4
PGN_ERR_NCCL = -5
PGN_ERR_STATE = -6
_ERR_NAMES = {
    -1: "PGN_ERR_INVALID",
    -2: "PGN_ERR_CUDA",
    -3: "PGN_ERR_NOMEM",
    -4: "PGN_ERR_NO_DEVICE",
    -5: "PGN_ERR_NCCL",
    -6: "PGN_ERR_STATE",
}

PGN_MEM_PARITY = 0
PGN_MEM_NATIVE = 1
PGN_ITERS_AFTER_SPLIT = 0
PGN_ITERS_BEFORE_CHECK = 1


class PaganiError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        name = _ERR_NAMES.get(code, str(code))
        super().__init__(f"{where} failed: {name}" + (f" ({detail})" if detail else ""))


class pgn_opts(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("mem_mode", C.c_int),
        ("max_iters", C.c_int),
        ("trace", C.c_int),
        ("trace_arrays", C.c_int),
        ("native_mem_fraction", C.c_double),
    ]


class pgn_result_t(C.Structure):
    _fields_ = [
        ("estimate", C.c_double),
        ("errorest", C.c_double),
        ("neval", C.c_uint64),
        ("nregions", C.c_uint64),
        ("nFinishedRegions", C.c_uint64),
        ("status", C.c_int32),
        ("lastPhase", C.c_int32),
        ("chi_sq", C.c_double),
        ("iters", C.c_uint64),
    ]


class pgn_rule_view(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32),
        ("feval", C.c_int32),
        ("point_code", C.c_void_p),
        ("weights", C.c_void_p),
        ("scale", C.c_void_p),
        ("norm", C.c_void_p),
        ("half_plus_g", C.c_double * PGN_NCODES),
        ("half_minus_g", C.c_double * PGN_NCODES),
        ("ratio", C.c_double),
        ("two_one_minus_ratio_unused", C.c_double),
        ("weights_host", C.c_double * (PGN_NSETS * PGN_NRULES)),
    ]


class pgn_heuristic_result(C.Structure):
    _fields_ = [
        ("success", C.c_int32),
        ("must_terminate", C.c_int32),
        ("probes", C.c_int32),
        ("reserved", C.c_int32),
        ("num_active", C.c_uint64),
        ("threshold", C.c_double),
        ("finished_errorest", C.c_double),
    ]


class pgn_iter_stats(C.Structure):
    _fields_ = [
        ("iteration", C.c_uint64),
        ("nregions", C.c_uint64),
        ("nregions_global", C.c_uint64),
        ("nactive", C.c_uint64),
        ("iter_estimate", C.c_double),
        ("iter_errorest", C.c_double),
        ("finished_estimate", C.c_double),
        ("finished_errorest", C.c_double),
        ("cum_estimate", C.c_double),
        ("cum_errorest", C.c_double),
        ("heuristic_ran", C.c_int32),
        ("heuristic_probes", C.c_int32),
        ("budget_overflow", C.c_int32),
        ("rebalanced", C.c_int32),
        ("threshold", C.c_double),
        ("dig_left", C.c_uint64),
        ("dig_length", C.c_uint64),
        ("dig_estimate", C.c_uint64),
        ("dig_errorest_raw", C.c_uint64),
        ("dig_errorest", C.c_uint64),
        ("dig_axis", C.c_uint64),
        ("dig_active", C.c_uint64),
        ("eval_ms", C.c_double),
        ("classify_ms", C.c_double),
        ("heuristic_ms", C.c_double),
        ("split_ms", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class pgn_integrate_args(C.Structure):
    _fields_ = [
        ("epsrel", C.c_double),
        ("epsabs", C.c_double),
        ("relerr_classification", C.c_int),
        ("use_current_regions", C.c_int),
        ("iters_mode", C.c_int),
        ("reserved", C.c_int),
    ]


pgn_eval_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
pgn_allreduce_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int)
pgn_alltoallv_fn = C.CFUNCTYPE(
    C.c_int,
    C.c_void_p,
    C.POINTER(C.c_double),
    C.POINTER(C.c_uint64),
    C.POINTER(C.c_uint64),
    C.POINTER(C.c_double),
    C.POINTER(C.c_uint64),
    C.POINTER(C.c_uint64),
)

_P = C.POINTER
_vp = C.c_void_p
_dp = _P(C.c_double)

# every symbol include/pagani_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pgn_create": (C.c_int, [_P(_vp), C.c_int, _P(pgn_opts)]),
    "pgn_destroy": (None, [_vp]),
    "pgn_last_error": (C.c_char_p, []),
    "pgn_version": (C.c_char_p, []),
    "pgn_device_info": (C.c_int, [_vp, _P(C.c_int), _P(C.c_size_t), _P(C.c_size_t)]),
    "pgn_set_eval_variant": (C.c_int, [_vp, C.c_int]),
    "pgn_get_stream": (C.c_int, [_vp, _P(_vp)]),
    "pgn_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "pgn_feval": (C.c_int, [C.c_int]),
    "pgn_rule_tables_host": (C.c_int, [C.c_int, _dp, _dp, _dp, _dp, _P(C.c_int32), _P(C.c_uint64)]),
    "pgn_rule_view_get": (C.c_int, [_vp, _P(pgn_rule_view)]),
    "pgn_regions_init_uniform": (C.c_int, [_vp, C.c_int]),
    "pgn_regions_import": (C.c_int, [_vp, _dp, _dp, C.c_size_t]),
    "pgn_regions_export": (C.c_int, [_vp, _dp, _dp, C.c_size_t, _P(C.c_size_t)]),
    "pgn_regions_count": (C.c_int, [_vp, _P(C.c_size_t)]),
    "pgn_set_volume": (C.c_int, [_vp, _dp, _dp]),
    "pgn_evaluate": (C.c_int, [_vp, _vp, _vp, _dp, _dp]),
    "pgn_classify": (C.c_int, [_vp, C.c_double, C.c_int, _P(pgn_iter_stats)]),
    "pgn_set_active": (C.c_int, [_vp, _P(C.c_uint8), C.c_size_t]),
    "pgn_set_estimates": (C.c_int, [_vp, _dp, _dp, _P(C.c_int32), C.c_size_t]),
    "pgn_heuristic_classify": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         _P(pgn_heuristic_result)]),
    "pgn_filter": (C.c_int, [_vp, _P(C.c_size_t)]),
    "pgn_split": (C.c_int, [_vp]),
    "pgn_fetch_estimates": (C.c_int, [_vp, _dp, _dp, _P(C.c_int32), _P(C.c_uint8), C.c_size_t]),
    "pgn_fetch_parents": (C.c_int, [_vp, _dp, _dp, _P(C.c_int32), C.c_size_t, _P(C.c_size_t)]),
    "pgn_integrate": (C.c_int, [_vp, _vp, _vp, _P(pgn_integrate_args), _P(pgn_result_t)]),
    "pgn_genz_available": (C.c_int, [C.c_int, C.c_int]),
    "pgn_genz_evaluator": (C.c_int, [C.c_int, C.c_int, _P(_vp), _P(_vp)]),
    "pgn_integrate_genz": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _dp, _dp, _P(pgn_result_t)]),
    "pgn_genz_true_value": (C.c_double, [C.c_int, C.c_int]),
    "pgn_genz_flops_per_eval": (C.c_double, [C.c_int, C.c_int]),
    "pgn_trace_count": (C.c_int, [_vp]),
    "pgn_trace_get": (C.c_int, [_vp, C.c_int, _P(pgn_iter_stats)]),
    "pgn_digest64": (C.c_uint64, [_vp, C.c_size_t, C.c_size_t]),
    "pgn_nccl_unique_id": (C.c_int, [_P(C.c_uint8)]),
    "pgn_comm_init": (C.c_int, [_vp, _P(C.c_uint8), C.c_int, C.c_int]),
    "pgn_comm_size": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int)]),
    "pgn_dev_uniform_split": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_size_t]),
    "pgn_dev_evaluate": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp, _dp, _dp]),
    "pgn_dev_count_active": (C.c_int, [_vp, _vp, C.c_size_t, _P(C.c_size_t)]),
    "pgn_dev_filter": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    "pgn_dev_split": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp]),
    "pgn_regions_import_device": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "pgn_regions_export_device": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _P(C.c_size_t)]),
    "pgn_set_rebalance": (C.c_int, [_vp, C.c_double, C.c_size_t]),
    "pgn_comm_set_host_collective": (C.c_int, [_vp, C.c_int, C.c_int, pgn_allreduce_fn, pgn_alltoallv_fn, _vp]),
    "pgn_plan_rebalance": (C.c_int, [_P(C.c_uint64), C.c_int, _P(C.c_uint64), _P(C.c_uint64), C.c_int, _P(C.c_int)]),
    "pgn_bench_fp64_peak": (C.c_int, [C.c_int, C.c_int, _dp]),
    "pgn_bench_hbm_copy": (C.c_int, [C.c_int, C.c_size_t, C.c_int, _dp]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpagani_b200.so")
_lib = None


def load():
    """Load libpagani_b200.so (built in-tree by `python -m gpuintegration_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -m gpuintegration_b200.build, or __graft_entry__.build()). "
            "pagani_b200 has no CPU / pure-Python fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, where):
    if rc != PGN_OK:
        detail = load().pgn_last_error()
        raise PaganiError(rc, where, detail.decode() if detail else "")
    return rc
