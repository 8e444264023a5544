"""Kernel-only driver for Nsight Compute: evaluate a uniform grid of parts^ndim regions `reps` times
(the reference's kernel-only profiling protocol, cuda/pagani/demos/new_time_and_call.cuh:30-70).
usage: python tools/ncu_eval.py family ndim parts [reps] [variant]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuintegration_b200 as pg  # noqa: E402

family, ndim, parts = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
with pg.Workspace(ndim) as ws:
    ws.set_eval_variant(variant)
    g = pg.Genz(family, ndim)
    for _ in range(reps):
        ws.uniform_split(parts)
        e, r = ws.apply_cubature_integration_rules(g)
    print(g.name, parts**ndim, "regions", e, r)
