"""BASELINE.json's five configurations, measured side by side on one box (GPU box only):
ours (B200-native engine) vs the unmodified reference CUDA PAGANI (oracle/_ref/ref_pagani) vs CPU Cuhre
(restated, all host cores, time-capped).  One JSON line per configuration; `python tools/baseline_configs.py
> profiles/rNN_configs.jsonl`.  Config 5's multi-GPU leg is bench.py / tests' job; here it runs on one GPU in
both memory models.

TEST/REPORT TOOLING: uses oracle/ only for the two reported baselines, never for our numbers."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpuintegration_b200 as pg  # noqa: E402
import oracle_cpu  # noqa: E402
import refharness  # noqa: E402

CONFIGS = [
    ("config1 5D Gaussian epsrel 1e-3", 4, 5, [1e-3], "parity"),
    ("config2 8D oscillatory sweep", 1, 8, [1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8], "parity"),
    ("config3 6D product peak epsrel 1e-7", 2, 6, [1e-7], "parity"),
    ("config4a 10D C0 (ours: exp(-10 sum|x-.5|))", 5, 10, [1e-3, 1e-4], "parity"),
    ("config4b 10D discontinuous (ours, SURVEY 8d parameters)", 6, 10, [1e-3, 1e-4], "parity"),
    ("config5 8D Gaussian epsrel 1e-9 (1 GPU, reference memory model)", 4, 8, [1e-9], "parity"),
    ("config5 8D Gaussian epsrel 1e-9 (1 GPU, native memory model)", 4, 8, [1e-9], "native"),
]
CPU_CAP_S = float(os.environ.get("CPU_CAP_S", "8"))


def ours(family, ndim, eps, mem_mode):
    g = pg.Genz(family, ndim)
    with pg.Workspace(ndim, mem_mode=mem_mode, trace=True) as ws:
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            r = ws.integrate(g, eps, 1e-40)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        tr = ws.trace()
    ev = sum(t["eval_ms"] for t in tr)
    return {"ms": best * 1e3, "estimate": r.estimate, "errorest": r.errorest, "status": r.status,
            "nregions": r.nregions, "iters": r.iters, "neval": r.neval, "evals_per_s": r.neval / best,
            "eval_kernel_share": ev / (best * 1e3), "max_regions": max(t["nregions_global"] for t in tr),
            "heuristic_iters": sum(1 for t in tr if t["heuristic_probes"] > 0),
            "true_value": g.true_value, "rel_err_true": abs(r.estimate - g.true_value) / abs(g.true_value),
            "roofline_frac_algorithmic": (sum(t["nregions"] for t in tr) * pg.feval(ndim) * g.flops_per_eval
                                          / (ev * 1e-3) / 1e12) if ev else None}


def main():
    peak = pg.fp64_peak_tflops(0)
    print(json.dumps({"fp64_peak_tflops_measured": peak, "host_cores": os.cpu_count()}), flush=True)
    for label, family, ndim, epsrels, mem_mode in CONFIGS:
        for eps in epsrels:
            rec = {"config": label, "family": family, "ndim": ndim, "epsrel": eps, "mem_mode": mem_mode}
            rec["ours"] = ours(family, ndim, eps, mem_mode)
            if rec["ours"]["roofline_frac_algorithmic"]:
                rec["ours"]["roofline_frac_algorithmic"] /= peak
            if mem_mode == "parity" and refharness.available():
                try:
                    ref = refharness.integrate(refharness.genz_name(family, ndim), eps, 1e-40, repeats=2)[-1]
                    rec["reference_gpu"] = {"ms": ref["ms"], "estimate": ref["estimate"], "errorest": ref["errorest"],
                                            "status": ref["status"], "nregions": ref["nregions"]}
                    rec["same_nregions_status"] = (ref["nregions"] == rec["ours"]["nregions"]
                                                   and ref["status"] == rec["ours"]["status"])
                    rec["estimate_diff_over_errorest"] = (abs(ref["estimate"] - rec["ours"]["estimate"])
                                                          / max(ref["errorest"] + rec["ours"]["errorest"], 1e-300))
                    rec["speedup_vs_reference_gpu"] = ref["ms"] / rec["ours"]["ms"]
                except Exception as e:  # noqa: BLE001
                    rec["reference_gpu"] = {"error": repr(e)[:200]}
            if mem_mode == "parity":
                try:
                    c = oracle_cpu.cuhre(family, ndim, eps, epsabs=1e-100, maxeval=10**11, threads=0, time_cap_s=CPU_CAP_S)
                    rec["cpu_cuhre_restated"] = {"seconds": c["seconds"], "status": c["status"], "neval": c["neval"],
                                                 "evals_per_s": c["neval"] / c["seconds"], "threads": c["threads"],
                                                 "rel_errorest_reached": c["errorest"] / abs(c["estimate"]) if c["estimate"] else None,
                                                 "cap_s": CPU_CAP_S}
                except Exception as e:  # noqa: BLE001
                    rec["cpu_cuhre_restated"] = {"error": repr(e)[:200]}
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
