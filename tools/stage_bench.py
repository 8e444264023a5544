"""Device time of the HBM-bound stages at the largest iterations of two trajectories (F_1_8D in the reference's
memory model): epsrel 1e-5 doubles the list to 16.8 M regions with every region surviving, epsrel 1e-4 has the
partially surviving iterations (about half the regions finished).  One JSON line per iteration with the algorithmic
roofline figures of SURVEY 8d (classify 29 B per region, compact+split 48 ndim + 36 B per survivor).
usage: python tools/stage_bench.py [label]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuintegration_b200 as pg  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else ""
peak = 6454.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
ndim = 8
with pg.Workspace(ndim, trace=True) as ws:
    g = pg.Genz(1, ndim)
    for eps in (1e-5, 1e-4):
        best = {}
        for _rep in range(3):
            ws.integrate(g, eps, 1e-40)
            for t in ws.trace():
                if t["nregions"] >= 2_000_000:
                    k = t["iteration"]
                    b = best.setdefault(k, dict(t))
                    for key in ("classify_ms", "split_ms", "heuristic_ms", "eval_ms"):
                        b[key] = min(b[key], t[key])
        for k, t in sorted(best.items()):
            n, m = t["nregions"], t["nactive"]
            cls = 29.0 * n / (t["classify_ms"] * 1e-3) / 1e9 if t["classify_ms"] else None
            spl = (48 * ndim + 36) * m / (t["split_ms"] * 1e-3) / 1e9 if t["split_ms"] else None
            print(json.dumps({"label": label, "epsrel": eps, "iteration": k, "regions": n, "survivors": m,
                              "classify_ms": t["classify_ms"], "classify_gbs": cls, "classify_frac": cls / peak if cls else None,
                              "split_ms": t["split_ms"], "compact_split_gbs": spl, "compact_split_frac": spl / peak if spl else None,
                              "heuristic_ms": t["heuristic_ms"], "eval_ms": t["eval_ms"]}), flush=True)
