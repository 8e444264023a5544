"""Compare the B200-native engine with the reference PAGANI iteration by iteration (GPU box).

    python tools/parity_sweep.py [--out gpurun_out/parity] [--quick]

For every configuration it runs oracle/_ref/ref_pagani in trace mode (per-iteration digests of the
region bounds, raw/refined errors, estimates, split axes and active flags) and our engine with
trace_arrays=True, and reports the first divergence, if any.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import gpuintegration_b200 as pg  # noqa: E402
import refharness  # noqa: E402

DIG = ["dig_left", "dig_length", "dig_estimate", "dig_errorest_raw", "dig_errorest", "dig_axis", "dig_active"]


def compare(family, ndim, epsrel, outdir, full=0):
    name = refharness.genz_name(family, ndim)
    prefix = os.path.join(outdir, f"{name}_{epsrel:g}")
    t0 = time.time()
    summary, ref_iters = refharness.trace(name, epsrel, prefix, full=full)
    t_ref = time.time() - t0
    ws = pg.Workspace(ndim, trace_arrays=True)
    t0 = time.time()
    res = ws.integrate(pg.Genz(family, ndim), epsrel, 1e-40)
    t_mine = time.time() - t0
    mine = ws.trace()
    ws.close()
    report = {
        "integrand": name, "epsrel": epsrel, "ref": summary, "ref_trace_s": t_ref, "mine_trace_s": t_mine,
        "mine": {"estimate": res.estimate, "errorest": res.errorest, "nregions": res.nregions,
                 "iters": res.iters, "status": res.status, "neval": res.neval},
        "iterations_ref": len(ref_iters), "iterations_mine": len(mine), "first_divergence": None,
    }
    for k, (r, m) in enumerate(zip(ref_iters, mine)):
        bad = []
        if r["n"] != m["nregions_global"]:
            bad.append(("n", r["n"], m["nregions_global"]))
        for d in DIG:
            if r[d] != m[d]:
                bad.append((d, r[d], m[d]))
        if r["outcome"] == "split" and r["nactive"] != m["nactive"]:
            bad.append(("nactive", r["nactive"], m["nactive"]))
        if bad:
            report["first_divergence"] = {"iteration": k, "fields": [b[0] for b in bad],
                                          "ref_iter": r, "mine_iter": m}
            break
    same_final = (summary["estimate"] == res.estimate and summary["errorest"] == res.errorest
                  and summary["nregions"] == res.nregions and summary["iters"] == res.iters
                  and summary["status"] == res.status)
    report["final_bit_identical"] = bool(same_final)
    report["final_rel_diff"] = abs(summary["estimate"] - res.estimate) / max(abs(summary["estimate"]), 1e-300)
    report["trajectory_identical"] = report["first_divergence"] is None and len(ref_iters) == len(mine)
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/parity")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--full", type=int, default=0)
    ap.add_argument("--cases", default="", help='explicit list "family,ndim,epsrel;family,ndim,epsrel;..."')
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.quick:
        cases = [(1, 8, 1e-3), (4, 5, 1e-3), (5, 8, 1e-3), (6, 6, 1e-3), (2, 6, 1e-3), (3, 8, 1e-3)]
    else:
        cases = [(1, 8, 1e-3), (1, 8, 1e-5), (1, 8, 1e-7), (4, 5, 1e-3), (4, 8, 1e-5), (5, 8, 1e-4),
                 (6, 6, 1e-3), (2, 6, 1e-5), (3, 8, 1e-3), (5, 10, 1e-3), (6, 10, 1e-3), (1, 5, 1e-6),
                 (2, 5, 1e-4), (3, 6, 1e-4), (4, 7, 1e-4), (6, 7, 1e-3)]
    if a.cases:
        cases = [(int(c.split(",")[0]), int(c.split(",")[1]), float(c.split(",")[2])) for c in a.cases.split(";") if c]
    ok = True
    with open(os.path.join(a.out, "parity_report.jsonl"), "w") as f:
        for fam, nd, eps in cases:
            try:
                r = compare(fam, nd, eps, a.out, full=a.full)
            except Exception as e:  # keep going: one report line per case
                r = {"integrand": refharness.genz_name(fam, nd), "epsrel": eps, "error": repr(e)}
            f.write(json.dumps(r) + "\n")
            f.flush()
            good = r.get("trajectory_identical") and r.get("final_bit_identical")
            ok = ok and bool(good)
            print(r.get("integrand"), eps, "trajectory_identical=", r.get("trajectory_identical"),
                  "final_bit_identical=", r.get("final_bit_identical"), "iters=", r.get("iterations_ref"),
                  "div=", (r.get("first_divergence") or {}).get("fields"), r.get("error", ""), flush=True)
    print("ALL PARITY OK" if ok else "PARITY FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
