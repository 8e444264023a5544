"""Machine-code parity of the Genz functors (include/pagani_b200/genz.cuh) with the reference's
(common/cuda/integrands.cuh:1303-1932), without a GPU.

Region-trajectory parity needs f(x) BIT-identical (symmetric integrands tie their fourth differences and
rounding decides the split axis).  Both functor sets are compiled by the same nvcc for sm_100a into
stand-alone kernels; the FP64 instructions of each pair must be the same multiset, where an operand that
comes straight from argument k (through the `- beta` subtraction) is written as `xk` -- this is what
distinguishes "round 9*x1, fuse 10*x0" from "round 10*x0, fuse 9*x1", a choice the compiler makes on its
own when two products feed one addition.

Runs only where the reference tree and nvcc exist (the build container); skipped elsewhere.
"""
import collections
import os
import re
import shutil
import subprocess

import pytest

REFERENCE = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = {1: "Oscillatory", 2: "ProductPeak", 3: "CornerPeak", 4: "Gaussian", 5: "CZero", 6: "Discontinuous"}
DIMS = (5, 6, 7, 8)

pytestmark = pytest.mark.skipif(
    not os.path.exists(os.path.join(REFERENCE, "common/cuda/integrands.cuh")) or shutil.which("nvcc") is None,
    reason="needs the reference tree and nvcc (build container only)")


def _signatures(sass):
    """{kernel: Counter of normalised FP64 instructions}"""
    out, cur, origin = {}, None, {}
    imm = re.compile(r"-?\d[0-9.e+]*|[+-]INF")

    def kill(dst, wide):
        m = re.fullmatch(r"R(\d+)", dst)
        if not m:
            return
        n = int(m.group(1))
        for r in (n, n - 1) + ((n + 1,) if wide else ()):
            origin.pop(f"R{r}", None)

    for line in sass.splitlines():
        m = re.search(r"Function : _Z\d+(k(?:ref|our)_\d_\d)PKdPd", line)
        if m:
            cur, origin = m.group(1), {}
            out[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m or cur is None:
            continue
        ins = m.group(1).strip()
        pred = ""
        if ins.startswith("@"):
            pred, ins = ins.split(None, 1)
        op = ins.split()[0]
        ops = [o.strip() for o in ins[len(op):].split(",")]
        base = op.split(".")[0]
        if base == "LDG":
            # LDG.E.64 Rd, desc[..][Ra.64+0x8]  -> Rd holds argument 1
            off = re.search(r"\+0x([0-9a-f]+)\]", ins)
            kill(ops[0], True)
            origin[ops[0]] = f"x{(int(off.group(1), 16) if off else 0) // 8}"
            continue
        if base not in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "F2I", "I2F"):
            if ops:
                kill(ops[0], "64" in op or "WIDE" in op)
            continue
        # operands are labelled with their argument only in a multiply(-add) by a literal constant:
        # that is the accumulation chain whose contraction the compiler is free to choose
        label = base in ("DFMA", "DMUL") and any(imm.fullmatch(o) for o in ops[1:])

        def name(o):
            core = o.replace(".reuse", "")
            reg = re.search(r"U?R\d+|RZ", core)
            if label and reg and reg.group(0) in origin:
                return core.replace(reg.group(0), origin[reg.group(0)])
            return re.sub(r"U?R\d+", "r", re.sub(r"U?P\d", "p", core))

        named = [name(o) for o in ops[1:]]
        # predication is not part of the signature: the compiler if-converts pow()'s special-case
        # paths differently around a loop than around one long expression, with identical values
        out[cur][f"{op} {', '.join(named)}"] += 1
        dst = ops[0]
        src = re.sub(r"[|-]", "", ops[1].replace(".reuse", "")) if len(ops) > 1 else ""
        keep = None
        # x_k - beta keeps the identity of argument k
        if base == "DADD" and len(ops) == 3 and src in origin and imm.fullmatch(ops[2]):
            keep = origin[src]
        kill(dst, True)
        if keep:
            origin[dst] = keep
    return out


@pytest.fixture(scope="module")
def signatures(tmp_path_factory):
    d = tmp_path_factory.mktemp("sass")
    src = ["#include <array>", '#include "common/cuda/integrands.cuh"', '#include "pagani_b200/genz.cuh"']
    for f, cls in FAMILIES.items():
        for n in DIMS:
            args = ",".join(f"x[{i}]" for i in range(n))
            src.append(f"__global__ void kref_{f}_{n}(const double* x, double* o)"
                       f"{{ F_{f}_{n}D f; o[threadIdx.x] = f({args}); }}")
            src.append(f"__global__ void kour_{f}_{n}(const double* x, double* o)"
                       f"{{ pagani_b200::genz::{cls}<{n}> f; o[threadIdx.x] = f({args}); }}")
    cu = d / "functors.cu"
    cu.write_text("\n".join(src))
    subprocess.check_call(["nvcc", "-std=c++17", "-O3", "--expt-relaxed-constexpr", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-I" + REFERENCE, "-I" + os.path.join(ROOT, "include"),
                           # the math-library spelling of the functors: fastmath.cuh's load-free cos is pinned
                           # against cos() bit for bit on the GPU instead (tests/test_gpu_fastmath.py)
                           "-DPGN_GENZ_LIBRARY_MATH",
                           "-w", "-cubin", "-o", str(d / "functors.cubin"), str(cu)])
    sass = subprocess.check_output(["cuobjdump", "-sass", str(d / "functors.cubin")], text=True)
    return _signatures(sass)


@pytest.mark.parametrize("family", sorted(FAMILIES))
@pytest.mark.parametrize("ndim", DIMS)
def test_functor_machine_code_matches_reference(signatures, family, ndim):
    ref, ours = signatures[f"kref_{family}_{ndim}"], signatures[f"kour_{family}_{ndim}"]
    assert sum(ref.values()) > 0
    assert ours == ref, {"only_reference": dict(ref - ours), "only_ours": dict(ours - ref)}
