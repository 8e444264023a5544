"""Per-probe log of the heuristic threshold search for one case, device-side and host-side search
(PGN_SEARCH_DEBUG=1, stderr).  usage: python tools/search_debug.py family ndim epsrel"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fam, nd, eps = sys.argv[1:4]
code = ("import sys; sys.path.insert(0, %r); import gpuintegration_b200 as pg\n"
        "ws = pg.Workspace(%s, trace=True); r = ws.integrate(pg.Genz(%s, %s), %s, 1e-40); print(r)\n"
        "print([(t['nregions_global'], t['nactive'], t['heuristic_probes']) for t in ws.trace() if t['heuristic_ran']])\n"
        % (ROOT, nd, fam, nd, eps))
for host in ("0", "1"):
    env = dict(os.environ, PGN_SEARCH_DEBUG="1", PGN_HOST_SEARCH=host)
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("==== PGN_HOST_SEARCH=" + host)
    print(p.stdout[-3000:])
    print(p.stderr[-20000:])
