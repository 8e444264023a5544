"""Quick GPU-box probe: roofline denominators and a first timing of ours vs the reference."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpuintegration_b200 as pg  # noqa: E402
import refharness  # noqa: E402

out = {}
out["fp64_peak_tflops"] = pg.fp64_peak_tflops()
out["hbm_copy_gbs"] = pg.hbm_copy_gbs()
print(json.dumps(out), flush=True)
cases = [(1, 8, 1e-3), (1, 8, 1e-5), (1, 8, 1e-6), (1, 8, 1e-7), (4, 8, 1e-5), (2, 6, 1e-5)]
if len(sys.argv) > 1:
    cases = [tuple(float(x) if i == 2 else int(x) for i, x in enumerate(c.split(","))) for c in sys.argv[1:]]
for fam, nd, eps in cases:
    ws = pg.Workspace(nd, trace=True)
    g = pg.Genz(fam, nd)
    times = []
    for rep in range(3):
        t0 = time.perf_counter()
        r = ws.integrate(g, eps, 1e-40)
        times.append(time.perf_counter() - t0)
    tr = ws.trace()
    ws.close()
    eval_ms = sum(t["eval_ms"] for t in tr)
    rec = {"integrand": g.name, "epsrel": eps, "estimate": r.estimate, "errorest": r.errorest, "status": r.status,
           "nregions": r.nregions, "iters": r.iters, "neval": r.neval, "ms": [t * 1e3 for t in times],
           "eval_ms": eval_ms, "classify_ms": sum(t["classify_ms"] for t in tr),
           "heur_ms": sum(t["heuristic_ms"] for t in tr), "split_ms": sum(t["split_ms"] for t in tr),
           "evals_per_s": r.neval / min(times), "eval_kernel_evals_per_s": r.neval / (eval_ms * 1e-3) if eval_ms else None,
           "true": g.true_value, "max_regions": max(t["nregions"] for t in tr)}
    if refharness.available():
        try:
            ref = refharness.integrate(refharness.genz_name(fam, nd), eps, repeats=3)
            rec["ref_ms"] = [l["ms"] for l in ref]
            rec["ref_estimate"] = ref[-1]["estimate"]
            rec["ref_nregions"] = ref[-1]["nregions"]
            rec["ref_status"] = ref[-1]["status"]
        except Exception as e:
            rec["ref_error"] = repr(e)
    print(json.dumps(rec), flush=True)
