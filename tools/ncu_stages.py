"""Driver for Nsight Compute captures of the HBM-bound stages: one integrate call whose last iterations
hold 8.4 M / 16.8 M regions (F_1_8D, epsrel 1e-5, parity memory model).
usage: ncu --set full -k regex:'k_classify|k_compact' --launch-skip 24 --launch-count 7 python tools/ncu_stages.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuintegration_b200 as pg  # noqa: E402

family, ndim, eps = 1, 8, 1e-5
if len(sys.argv) > 3:
    family, ndim, eps = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
with pg.Workspace(ndim, trace=True) as ws:
    r = ws.integrate(pg.Genz(family, ndim), eps, 1e-40)
    for t in ws.trace()[-4:]:
        print({k: t[k] for k in ("nregions", "nactive", "classify_ms", "split_ms", "eval_ms")})
    print(r)
