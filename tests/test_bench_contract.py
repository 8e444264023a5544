"""bench.py's JSON line against the driver's contract, without a GPU: the engine, torch.cuda and the side legs are
replaced by stand-ins that return plausible per-iteration statistics, so that a slip in the (GPU-only) result
assembly -- a missing key, an undefined variable -- shows up in the CPU suite and not at round end."""
import importlib
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Result:
    def __init__(self, neval, nregions, status):
        self.estimate, self.errorest = 3.4395579e-05, 4.4e-09
        self.neval, self.nregions, self.status, self.iters = neval, nregions, status, 3


class _Workspace:
    launch_count = 0
    cuda_stream = 0

    def __init__(self, ndim, **kw):
        self.ndim = ndim

    def integrate(self, g, eps, epsabs, subregions=False):
        _Workspace.launch_count += 9
        self._tr = []
        for it, (n, act) in enumerate([(256, 256), (512, 512), (1024, 492)]):
            self._tr.append({"iteration": it, "nregions": n, "nregions_global": n, "nactive": act, "eval_ms": 0.5 * (it + 1),
                             "classify_ms": 0.01, "heuristic_ms": 0.02 * it, "split_ms": 0.03, "heuristic_probes": it,
                             "heuristic_ran": int(it == 2), "budget_overflow": 0, "iter_estimate": 3.4e-5,
                             "iter_errorest": 1e-7})
        return _Result(sum(t["nregions"] for t in self._tr) * 1105, 1792, 0)

    def trace(self):
        return self._tr

    def uniform_split(self, parts):
        pass

    def export_regions(self):
        return np.zeros((self.ndim, 256)), np.full((self.ndim, 256), 0.5)

    def import_regions(self, left, length):
        assert left.shape == (self.ndim, 256)

    def close(self):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 12.5


@pytest.fixture
def bench(monkeypatch):
    import torch

    sys.path.insert(0, ROOT)
    import gpuintegration_b200 as pg
    from gpuintegration_b200 import parallel

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda ptr, device=None: object())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: types.SimpleNamespace(multi_processor_count=148))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(pg, "Workspace", _Workspace)
    monkeypatch.setattr(pg, "Genz", lambda f, n: types.SimpleNamespace(flops_per_eval=107.0, true_value=3.44e-05, name="F_1_8D"))
    monkeypatch.setattr(pg, "fp64_peak_tflops", lambda dev=0: 34.0)
    monkeypatch.setattr(parallel, "max_over_ranks", lambda v, group=None: v)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    sys.modules.pop("bench", None)
    mod = importlib.import_module("bench")
    import oracle_cpu
    import refharness

    monkeypatch.setattr(refharness, "available", lambda: False)
    monkeypatch.setattr(oracle_cpu, "cuhre", lambda *a, **k: {"neval": 10**8, "seconds": 1.0, "threads": 4, "status": 1,
                                                              "errorest": 1e-7, "estimate": 3.4e-5})
    return mod


CONTRACT = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]


def test_bench_line_carries_every_contract_key(bench, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "1"])
    real_stop = bench.ClockSampler.stop

    def stop(self):  # no nvidia-smi in the CPU container: keep the sampler's shutdown, supply the clocks
        real_stop(self)
        return {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}

    monkeypatch.setattr(bench.ClockSampler, "stop", stop)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    for k in CONTRACT:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "evals/s" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert d["roofline"]["frac"] == pytest.approx(d["roofline"]["achieved"] / d["roofline"]["peak"])
    assert d["roofline"]["traffic"] > 0
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert d["gpu_launches"] > 0
    # evaluations of one step / device time of the timed steps
    evals_per_step = 6 * (256 + 512 + 1024) * 1105
    assert d["value"] == pytest.approx(evals_per_step * 2 / 12.5e-3)
    # trajectory parity against the committed goldens is part of the line (the stand-in engine does not follow them)
    assert d["parity_ok"] is False and len(d["parity"]["cases"]) == 6
    assert all(c["case"].startswith("f1_8_") and c["ok"] is False for c in d["parity"]["cases"])
    # BASELINE configs 3-5 and the native-mode tolerance ladder ride along as side records
    assert [c["family"] for c in d["configs"]] == [2, 5, 6, 4] and all("parity" in c and "ours" in c for c in d["configs"])
    assert d["native_ladder"]["steps"] and d["native_ladder"]["steps"][0]["epsrel"] == 1e-3
    assert d["roofline"]["fp64_pipe_pct"] > 0 and d["roofline"]["frac_executed"] > 0
    ib = d["roofline"]["issue_bound"]  # from the committed ncu instruction mix and the sampled SM clock
    assert ib["evals_per_s"] == pytest.approx(148 * 4 * 1965e6 * 32 / ib["issue_cycles_per_32_evals"])
    assert ib["frac"] == pytest.approx(d["roofline"]["kernel_evals_per_s"] / ib["evals_per_s"])
    hbm = d["roofline_hbm"]
    assert "error" not in hbm and hbm["regions"] == 1024 and hbm["largest_all_active_iteration"]["regions"] == 512


def test_reference_arm_line(bench, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    d = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] == pytest.approx(1e8)
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 4
