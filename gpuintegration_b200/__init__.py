"""pagani_b200: B200-native (sm_100a) PAGANI adaptive cubature behind the reference's integrate API.

The compute path is the in-tree CUDA library gpuintegration_b200/libpagani_b200.so; importing the
package does not need a GPU (rule tables and the rebalance planner are host code), but every
evaluation does -- there is no CPU fallback.
"""
from . import _capi
from ._capi import PaganiError
from .pagani import (
    GENZ_NAMES,
    Genz,
    IntegrationResult,
    Volume,
    Workspace,
    digest64,
    feval,
    fp64_peak_tflops,
    hbm_copy_gbs,
    nccl_unique_id,
    plan_rebalance,
    rule_tables,
)

__all__ = [
    "GENZ_NAMES",
    "Genz",
    "IntegrationResult",
    "PaganiError",
    "Volume",
    "Workspace",
    "digest64",
    "feval",
    "fp64_peak_tflops",
    "hbm_copy_gbs",
    "nccl_unique_id",
    "plan_rebalance",
    "rule_tables",
]
