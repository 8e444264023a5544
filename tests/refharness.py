"""Helpers around oracle/_ref/ref_pagani (the unmodified reference PAGANI, built by oracle/Makefile).

TEST INFRASTRUCTURE: imported by tests/, tools/ and bench.py's reference-GPU side leg only.
"""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_pagani")

# harness integrand name -> (Genz family, ndim)
def genz_name(family, ndim):
    return f"f{family}_{ndim}"


def available():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run(args, timeout=1800):
    p = subprocess.run([REF_BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    return p.returncode, lines, p.stderr


def integrate(name, epsrel, epsabs=1e-40, repeats=3, plain=False, timeout=1800):
    args = ["integrate", name, repr(float(epsrel)), "--epsabs", repr(float(epsabs)), "--repeats", repeats]
    if plain:
        args.append("--plain")
    rc, lines, err = run(args, timeout)
    if rc != 0:
        raise RuntimeError(f"ref_pagani integrate {name} failed rc={rc}: {err[-500:]}")
    return lines


def trace(name, epsrel, out_prefix, epsabs=1e-40, full=0, max_iters=700, timeout=3600):
    """Run the re-assembled reference loop; returns (summary, [per-iteration dicts])."""
    os.makedirs(os.path.dirname(out_prefix) or ".", exist_ok=True)
    args = ["trace", name, repr(float(epsrel)), "--epsabs", repr(float(epsabs)), "--out", out_prefix,
            "--full", full, "--max-iters", max_iters]
    rc, lines, err = run(args, timeout)
    if rc not in (0, 3):
        raise RuntimeError(f"ref_pagani trace {name} failed rc={rc}: {err[-500:]}")
    iters = [json.loads(l) for l in open(out_prefix + ".jsonl")]
    for it in iters:
        for k in list(it):
            if k.startswith("dig_"):
                it[k] = int(it[k])
    return lines[-1], iters


def load_full(out_prefix, it):
    """Raw arrays of one iteration written by trace(..., full=N)."""
    raw = open(f"{out_prefix}.it{it}.bin", "rb").read()
    magic, ndim, n, itn = np.frombuffer(raw, dtype=np.uint64, count=4)
    assert magic == 0x50474E5245463031 and itn == it
    ndim, n = int(ndim), int(n)
    off = 32

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a

    d = {
        "left": take(np.float64, ndim * n).reshape(ndim, n),
        "length": take(np.float64, ndim * n).reshape(ndim, n),
        "estimate": take(np.float64, n),
        "errorest_raw": take(np.float64, n),
        "errorest": take(np.float64, n),
        "axis": take(np.int32, n),
        "active_relerr": take(np.uint8, n),
        "active": take(np.uint8, n),
    }
    assert off == len(raw)
    return d
