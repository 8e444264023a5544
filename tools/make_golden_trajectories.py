"""Record golden trajectories from the UNMODIFIED reference PAGANI (GPU box only).

    python tools/make_golden_trajectories.py [--out gpurun_out/golden]

Runs oracle/_ref/ref_pagani in trace mode for a fixed list of cases and writes
  <out>/trajectories/<integrand>_<epsrel>.json   per-iteration scalars + 64-bit digests + summary
  <out>/arrays/<integrand>_<epsrel>.npz          raw per-region arrays of the small iterations
The files are then copied (by hand) into tests/golden/, where the CPU and GPU suites read them.
TEST INFRASTRUCTURE.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import refharness  # noqa: E402

# (family, ndim, epsrel, keep raw arrays of iterations with <= this many regions)
CASES = [
    # BASELINE configs[1]: F_1_8D at all six tolerances of the bench step
    (1, 8, 1e-3, 1024), (1, 8, 1e-4, 0), (1, 8, 1e-5, 0), (1, 8, 1e-6, 0), (1, 8, 1e-7, 0), (1, 8, 1e-8, 0),
    # configs[0] integrand, configs[2] (F_2_6D 1e-7), configs[3] (10-D C0 / discontinuous), configs[4] (F_4_8D 1e-9)
    (4, 5, 1e-3, 1024), (2, 6, 1e-7, 0), (5, 10, 1e-3, 1024), (5, 10, 1e-4, 0), (6, 10, 1e-3, 1024), (6, 10, 1e-4, 0),
    (4, 8, 1e-9, 0),
    # the other families / dimensions
    (4, 8, 1e-4, 0), (5, 8, 1e-3, 1024), (6, 6, 1e-3, 1024), (2, 6, 1e-3, 512), (2, 6, 1e-5, 0), (3, 8, 1e-3, 1024),
    (1, 5, 1e-6, 512), (3, 5, 1e-5, 512), (4, 7, 1e-4, 0), (6, 7, 1e-3, 512), (0, 2, 1e-3, 1024), (7, 2, 1e-6, 1024),
    # multilinear integrands at 3 and 4 axes: split axes decided by rounding noise (the reference's contraction choices)
    (7, 3, 1e-8, 1024), (7, 4, 1e-8, 1024),
]
NAMES = {0: "const", 7: "xy"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    a = ap.parse_args()
    os.makedirs(os.path.join(a.out, "trajectories"), exist_ok=True)
    os.makedirs(os.path.join(a.out, "arrays"), exist_ok=True)
    os.makedirs(os.path.join(a.out, "scratch"), exist_ok=True)
    for fam, nd, eps, full in CASES:
        name = f"{NAMES[fam]}_{nd}" if fam in NAMES else refharness.genz_name(fam, nd)
        epsabs = 1e-12 if fam in NAMES else 1e-40
        tag = f"{name}_{eps:g}"
        prefix = os.path.join(a.out, "scratch", tag)
        summary, iters = refharness.trace(name, eps, prefix, epsabs=epsabs, full=full)
        for it in iters:
            for k in list(it):
                if k.startswith("dig_"):
                    it[k] = str(it[k])  # JSON has no 64-bit integers
        doc = {"integrand": name, "family": fam, "ndim": nd, "epsrel": eps, "epsabs": epsabs,
               "summary": summary, "iterations": iters}
        with open(os.path.join(a.out, "trajectories", tag + ".json"), "w") as f:
            json.dump(doc, f, indent=0)
        arrays = {}
        for it in iters:
            if it["full"]:
                d = refharness.load_full(prefix, it["it"])
                for k, v in d.items():
                    arrays[f"it{it['it']}_{k}"] = v
        if arrays:
            np.savez_compressed(os.path.join(a.out, "arrays", tag + ".npz"), **arrays)
        print(tag, "iterations", len(iters), "status", summary["status"], "matches_workspace",
              summary["matches_workspace"], flush=True)


if __name__ == "__main__":
    main()
