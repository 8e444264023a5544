"""the reference's benchmark/CSV driver (clean_time_and_call, genz_integrals.cu) on the
drop-in header.  Runs the 5-D slice of the Genz sweep at epsrel 1e-3 and 1e-4 and checks the CSV the
reference's analysis scripts would read: column layout, one row per timed run, converged estimates within
the requested tolerance of the closed-form value, and agreement with the Python front end."""
import csv
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tools", "demos", "_build", "genz_integrals")


@pytest.mark.gpu
def test_genz_sweep_csv(tmp_path):
    import gpuintegration_b200 as pg

    assert os.path.exists(EXE), "tools/demos/_build/genz_integrals missing: run __graft_entry__.build()"
    out = tmp_path / "genz.csv"
    p = subprocess.run([EXE, str(out), "4", "5"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    rows = list(csv.reader(open(out)))
    header = [h.strip() for h in rows[0]]
    assert header == ["id", "ndim", "use_custom", "integral", "epsrel", "epsabs", "estimate", "errorest",
                      "nregions", "nFinishedRegions", "iters", "status", "time"]
    body = rows[1:]
    assert len(body) == 6 * 2 * 2  # six families x two tolerances x (3 runs - 1 warm-up)
    seen = set()
    for r in body:
        assert len(r) == 13 and r[1] == "5" and r[2] == "library"
        fam, eps = int(r[0][1]), float(r[4])
        true, est, err, status = float(r[3]), float(r[6]), float(r[7]), int(r[11])
        assert float(r[5]) == 1e-40 and float(r[12]) > 0.0
        if status == 0:
            assert err <= eps * abs(est) * (1 + 1e-12)
            assert abs(est - true) <= 5 * eps * abs(true)
        if (fam, eps) not in seen:
            seen.add((fam, eps))
            with pg.Workspace(5) as ws:
                ref = ws.integrate(pg.Genz(fam, 5), eps, 1e-40)
            assert (ref.estimate, ref.errorest, ref.nregions, ref.status) == (est, err, int(r[8]), status) or (
                abs(ref.estimate - est) <= 1e-14 * abs(est) and ref.nregions == int(r[8]))
    assert len(seen) == 12
