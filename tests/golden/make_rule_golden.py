"""Regenerate tests/golden/rule_tables.npz from the reference's own host code.

TEST INFRASTRUCTURE.  Needs /root/reference (build container only):
    make -C oracle ref_rule_dump && python tests/golden/make_rule_golden.py
The tool oracle/_ref/ref_rule_dump runs quad::Rule<double>::Init
(/root/reference/cuda/pagani/quad/GPUquad/Rule.cuh:370-510) on the host and writes raw tables.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NDIMS = [2, 3, 4, 5, 6, 7, 8, 9, 10]
# 11..15 (PGN_MAX_NDIM) go into a second file: only the arrays the tests compare (the generator matrix of 15-D is
# 4.5 MB raw, but it holds 11 distinct values and compresses to a few per cent)
NDIMS_HI = [11, 12, 13, 14, 15]


def read_one(path):
    raw = open(path, "rb").read()
    ndim, feval, nsets, npos = np.frombuffer(raw, dtype=np.int32, count=4)
    off = 16

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off).copy()
        off += a.nbytes
        return a

    out = {
        "feval": np.int64(feval),
        "G": take(np.float64, ndim * nsets).reshape(nsets, ndim),
        "weights": take(np.float64, nsets * 5).reshape(nsets, 5),
        "scale": take(np.float64, nsets * 5).reshape(nsets, 5),
        "norm": take(np.float64, nsets * 5).reshape(nsets, 5),
        "set_of_point": take(np.int32, feval),
        "var_start": take(np.int32, feval + 1),
        "gen_pos": take(np.int32, npos),
        "generators": take(np.float64, ndim * feval).reshape(ndim, feval),
    }
    assert off == len(raw)
    # rules 0 and 4 of scale/norm are never written by the reference (malloc garbage)
    out["scale"][:, [0, 4]] = 0.0
    out["norm"][:, [0, 4]] = 0.0
    return out


def main():
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_rule_dump")
    if not os.path.exists(tool):
        sys.exit("build oracle/_ref/ref_rule_dump first (make -C oracle ref_rule_dump)")
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([tool, tmp] + [str(n) for n in NDIMS])
        for n in NDIMS:
            for k, v in read_one(os.path.join(tmp, f"rule_ndim{n}.bin")).items():
                arrays[f"n{n}_{k}"] = v
    dst = os.path.join(ROOT, "tests", "golden", "rule_tables.npz")
    np.savez_compressed(dst, **arrays)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([tool, tmp] + [str(n) for n in NDIMS_HI])
        for n in NDIMS_HI:
            one = read_one(os.path.join(tmp, f"rule_ndim{n}.bin"))
            for k in ("feval", "weights", "scale", "norm", "set_of_point", "generators"):
                arrays[f"n{n}_{k}"] = one[k]
    dst = os.path.join(ROOT, "tests", "golden", "rule_tables_hi.npz")
    np.savez_compressed(dst, **arrays)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
