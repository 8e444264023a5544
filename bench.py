"""Benchmark of the PAGANI hot path (BASELINE.json: integrand evals/sec and time-to-epsrel, 8-D Genz).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the workload BASELINE.json quotes the metric on (configs[1]): the 8-D Genz
oscillatory integrand F_1_8D integrated with PAGANI at epsrel 1e-3, 1e-4, ... 1e-8 (epsabs 1e-40,
unit cube, the reference's 16e9-byte memory model so the region trajectory is the reference's).
`value` = integrand evaluations of all ranks / device time (CUDA events on the engine's stream,
max over ranks); `e2e` = the same through the public API with the initial region list coming from
pinned host memory and results read back to the host, on the wall clock.

--impl reference times the reference's CPU path for this metric -- Cuhre, restated (libcuba is not
vendored in the reference tree; see oracle/cuhre_restated.cc) -- on all host cores, on a bounded
sample of the same integrand.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# the contract is ONE JSON line on stdout: NCCL's own banner / debug lines go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FAMILY, NDIM = 1, 8
EPSRELS = [1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8]
EPSABS = 1e-40
METRIC = "integrand evals/sec (PAGANI, 8-D Genz oscillatory F_1_8D, epsrel sweep 1e-3..1e-8)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [s for s in sm if mx and s > 0.5 * mx[0]] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank):
    """The reference arm: CPU Cuhre (restated) on the box's host cores, all threads."""
    if rank != 0:
        return
    import oracle_cpu

    cores = os.cpu_count() or 1
    cap = 10.0
    evals = seconds = 0.0
    last = None
    for i in range(args.warmup + args.steps):
        # threads given explicitly: under torchrun OMP_NUM_THREADS is 1, and "all threads" (0) would mean one
        r = oracle_cpu.cuhre(FAMILY, NDIM, EPSRELS[0], epsabs=1e-100, maxeval=10**11, threads=cores, time_cap_s=cap)
        if i >= args.warmup:
            evals += r["neval"]
            seconds += r["seconds"]
            last = r
    value = evals / seconds
    sample = (f"Cuhre (restated, not libcuba) on F_1_8D at epsrel {EPSRELS[0]:g}, each step capped at {cap:g} s "
              f"of wall time (status {last['status']}, relative error estimate reached "
              f"{last['errorest'] / abs(last['estimate']):.2e})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "F_1_8D (8-D Genz oscillatory), unit cube, epsabs 1e-100, CPU Cuhre bounded sample"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": last["threads"], "kind": "port",
                         "sample": sample, "host_cores_available": cores},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "trajectories")
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_eval_f1_8d.json")


def load_ncu_summary():
    """Figures of the committed `ncu --set full` capture of the evaluation kernel (written by tools/ncu_summary.py
    from the .ncu-rep): FP64 pipe utilisation, executed FP64 flops per evaluation, DRAM bytes per region."""
    try:
        return json.load(open(NCU_SUMMARY))
    except (OSError, ValueError):
        return None



def golden_tag(family, ndim, eps):
    return f"f{family}_{ndim}_{eps:g}"


def check_against_golden(family, ndim, eps, res, trace):
    """Compare one integrate() call with the trajectory the UNMODIFIED reference produced on a B200 for the same
    integrand and tolerance (tests/golden/trajectories, recorded by tools/make_golden_trajectories.py from
    oracle/_ref/ref_pagani): per-iteration global region count, surviving regions, heuristic-filter decisions,
    budget-overflow decisions, and the summary.  Works at any rank count (the quantities are global)."""
    path = os.path.join(GOLDEN_DIR, golden_tag(family, ndim, eps) + ".json")
    if not os.path.exists(path):
        return {"case": golden_tag(family, ndim, eps), "ok": None, "why": "no golden fixture"}
    doc = json.load(open(path))
    bad = []
    its, s = doc["iterations"], doc["summary"]
    if len(trace) != len(its):
        bad.append(f"iterations {len(trace)} != {len(its)}")
    for r, m in zip(its, trace):
        if m["nregions_global"] != r["n"]:
            bad.append(f"it {r['it']}: regions {m['nregions_global']} != {r['n']}")
        if m["heuristic_ran"] != r["heuristic_ran"]:
            bad.append(f"it {r['it']}: heuristic_ran {m['heuristic_ran']} != {r['heuristic_ran']}")
        if m["budget_overflow"] != r["budget_overflow"]:
            bad.append(f"it {r['it']}: budget_overflow {m['budget_overflow']} != {r['budget_overflow']}")
        if r["outcome"] == "split" and m["nactive"] != r["nactive"]:
            bad.append(f"it {r['it']}: active {m['nactive']} != {r['nactive']}")
        scale = max(abs(r["iter_estimate"]), r["iter_errorest"])
        if abs(m["iter_estimate"] - r["iter_estimate"]) > 1e-12 * scale:
            bad.append(f"it {r['it']}: iteration estimate differs by more than 1e-12")
    if (res.status, res.nregions, res.iters) != (s["status"], s["nregions"], s["iters"]):
        bad.append(f"summary (status, nregions, iters) {(res.status, res.nregions, res.iters)} != "
                   f"{(s['status'], s['nregions'], s['iters'])}")
    if abs(res.estimate - s["estimate"]) > 1e-12 * max(abs(s["estimate"]), s["errorest"]):
        bad.append("final estimate differs by more than 1e-12")
    return {"case": golden_tag(family, ndim, eps), "ok": not bad, "iterations": len(its), "mismatches": bad[:5]}


def timed_integrate(ws, g, eps, repeats=2):
    """min wall ms over `repeats` calls of ws.integrate (collective under torchrun) -> (ms, result, trace)."""
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = ws.integrate(g, eps, EPSABS)
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None or dt < best else best
    return best, r, ws.trace()


def reference_gpu_ms(name, eps, repeats=4):
    """Unmodified reference CUDA PAGANI on this GPU: min wall ms of Workspace::integrate over the repeats after the
    first (the reference's own protocol discards the first run, new_time_and_call.cuh:230-238)."""
    import refharness

    if not refharness.available():
        return None
    lines = refharness.integrate(name, eps, EPSABS, repeats=repeats)
    runs = [l for l in lines if "ms" in l]
    keep = runs[1:] if len(runs) > 1 else runs
    best = min(keep, key=lambda l: l["ms"])
    return {"ms": best["ms"], "ms_all": [l["ms"] for l in runs], "status": best["status"], "nregions": best["nregions"],
            "estimate": best.get("estimate"), "errorest": best.get("errorest")}


def cpu_cuhre(family, ndim, eps, cap_s, cores):
    import oracle_cpu

    c = oracle_cpu.cuhre(family, ndim, eps, epsabs=1e-100, maxeval=10**11, threads=cores, time_cap_s=cap_s)
    return {"seconds": c["seconds"], "status": c["status"], "neval": c["neval"], "threads": c["threads"],
            "evals_per_s": c["neval"] / c["seconds"] if c["seconds"] > 0 else None,
            "reached_relerr": c["errorest"] / abs(c["estimate"]) if c["estimate"] else None, "cap_s": cap_s,
            "label": "Cuhre (restated, not libcuba)"}


# BASELINE.json configs[2..4] as side records: (label, family, ndim, epsrel)
SIDE_CONFIGS = [
    ("config3 F_2_6D product peak, epsrel 1e-7", 2, 6, 1e-7),
    ("config4a F_5_10D C0, epsrel 1e-3", 5, 10, 1e-3),
    ("config4b F_6_10D discontinuous (SURVEY 8d parameters), epsrel 1e-3", 6, 10, 1e-3),
    ("config5 F_4_8D Gaussian, epsrel 1e-9", 4, 8, 1e-9),
]


def run_side_configs(pg, parallel, local_rank, world):
    """configs[2..4] in the reference's memory model, trajectory checked against the golden.  Collective: every rank
    calls it (add_side_baselines adds the reference kernels and CPU Cuhre on rank 0 afterwards)."""
    out = []
    for label, family, ndim, eps in SIDE_CONFIGS:
        ws = pg.Workspace(ndim, device=local_rank, mem_mode="parity", trace=True)
        if world > 1:
            parallel.attach_nccl(ws)
        g = pg.Genz(family, ndim)
        ms, r, tr = timed_integrate(ws, g, eps)
        ms = parallel.max_over_ranks(ms)
        ws.close()
        rec = {"config": label, "family": family, "ndim": ndim, "epsrel": eps, "n_gpus": world,
               "ours": {"ms": ms, "status": r.status, "estimate": r.estimate, "errorest": r.errorest, "nregions": r.nregions,
                        "iters": r.iters, "neval": r.neval, "evals_per_s": r.neval / (ms * 1e-3),
                        "max_regions_in_flight": max(t["nregions_global"] for t in tr),
                        "heuristic_iterations": sum(1 for t in tr if t["heuristic_ran"]),
                        "true_value": g.true_value, "rel_err_vs_true": abs(r.estimate - g.true_value) / abs(g.true_value)},
               "parity": check_against_golden(family, ndim, eps, r, tr)}
        out.append(rec)
    return out


def add_side_baselines(records, cores):
    """Rank 0 only, after every collective leg: the reference kernels on this GPU and CPU Cuhre at the same tolerance."""
    for rec in records:
        family, ndim, eps = rec["family"], rec["ndim"], rec["epsrel"]
        try:
            ref = reference_gpu_ms(f"f{family}_{ndim}", eps)
            if ref:
                rec["reference_gpu"] = ref
                rec["speedup_vs_reference_gpu"] = ref["ms"] / rec["ours"]["ms"]
                rec["reference_gpu"]["same_nregions_as_ours"] = ref["nregions"] == rec["ours"]["nregions"]
        except Exception as e:  # noqa: BLE001
            rec["reference_gpu"] = {"error": repr(e)}
        try:
            rec["cpu_cuhre_same_epsrel"] = cpu_cuhre(family, ndim, eps, 6.0, cores)
        except Exception as e:  # noqa: BLE001
            rec["cpu_cuhre_same_epsrel"] = {"error": repr(e)}


def run_native_ladder(pg, parallel, local_rank, world, budget_s):
    """The reference's tolerance ladder (new_interface_Genz1_8D.cu:9-25: epsrel divided by 5 until the first run
    that does not converge) on F_1_8D with the NATIVE memory model (free HBM of all ranks instead of the reference's
    hard-wired 16e9 bytes), so that tight tolerances are actually reached.  Stops at the first failure or when the
    wall-time budget is used up.  Collective."""
    ws = pg.Workspace(NDIM, device=local_rank, mem_mode="native", trace=True)
    if world > 1:
        parallel.attach_nccl(ws)
    g = pg.Genz(FAMILY, NDIM)
    steps, eps, t_start, stopped = [], 1e-3, time.perf_counter(), "first failure"
    while eps > 1e-12:
        t0 = time.perf_counter()
        r = ws.integrate(g, eps, EPSABS)
        ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3)
        tr = ws.trace()
        steps.append({"epsrel": eps, "status": r.status, "time_to_epsrel_ms": ms if r.status == 0 else None, "ms": ms,
                      "estimate": r.estimate, "errorest": r.errorest, "nregions": r.nregions, "iters": r.iters,
                      "neval": r.neval, "evals_per_s": r.neval / (ms * 1e-3),
                      "rel_err_vs_true": abs(r.estimate - g.true_value) / abs(g.true_value),
                      "max_regions_in_flight": max(t["nregions_global"] for t in tr)})
        if r.status != 0:
            break
        # every rank must take the same decision: the slowest rank's clock decides
        if parallel.max_over_ranks(time.perf_counter() - t_start) > budget_s:
            stopped = f"wall-time budget of {budget_s:g} s"
            break
        eps /= 5.0
    ws.close()
    reached = [s["epsrel"] for s in steps if s["status"] == 0]
    return {"integrand": "F_1_8D", "mem_mode": "native", "n_gpus": world, "stopped_by": stopped,
            "tightest_epsrel_reached": min(reached) if reached else None, "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mem-mode", default="parity", choices=["parity", "native"])
    ap.add_argument("--no-side-legs", action="store_true", help="skip cpu_baseline / reference-GPU / configs / ladder legs")
    ap.add_argument("--ladder-budget-s", type=float, default=45.0, help="wall-time budget of the native-mode tolerance ladder")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import numpy as np
    import torch

    import gpuintegration_b200 as pg
    from gpuintegration_b200 import parallel

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        parallel.init_process_group("nccl")
    ws = pg.Workspace(NDIM, device=local_rank, mem_mode=args.mem_mode, trace=True)
    if world > 1:
        parallel.attach_nccl(ws)
    g = pg.Genz(FAMILY, NDIM)
    stream = torch.cuda.ExternalStream(ws.cuda_stream, device=local_rank)
    feval = pg.feval(NDIM)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step(collect=None):
        for eps in EPSRELS:
            r = ws.integrate(g, eps, EPSABS)
            if collect is not None:
                collect.append((eps, r, ws.trace()))

    for _ in range(args.warmup):
        step()

    # ---- timed region: device time by CUDA events on the stream the kernels are launched on
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ws.launch_count
    collected = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step(collected)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    launches = ws.launch_count - launches0
    dev_s = parallel.max_over_ranks(dev_s)
    wall = parallel.max_over_ranks(wall)

    evals_per_step = sum(r.neval for _e, r, _t in collected[:len(EPSRELS)])
    total_evals = evals_per_step * args.steps
    value = total_evals / dev_s

    # ---- roofline of the dominant kernel (region evaluation): algorithmic flops / its measured time
    eval_ms = sum(t["eval_ms"] for _e, _r, tr in collected for t in tr)
    local_evals = sum(t["nregions"] for _e, _r, tr in collected for t in tr) * feval
    other_ms = {k: sum(t[k] for _e, _r, tr in collected for t in tr) for k in ("classify_ms", "heuristic_ms", "split_ms")}
    flops_per_eval = g.flops_per_eval
    achieved_tf = local_evals * flops_per_eval / (eval_ms * 1e-3) / 1e12
    n_eval_launches = sum(1 for _e, _r, tr in collected for t in tr if t["nregions"] > 0)

    # ---- end to end through the public API: initial regions from pinned host memory, results read back
    ws.uniform_split(2)
    left, length = ws.export_regions()
    n0 = left.shape[1]
    pin_left = torch.from_numpy(left).pin_memory()
    pin_len = torch.from_numpy(length).pin_memory()
    h2d = d2h = 0
    barrier()
    t0 = time.perf_counter()
    e2e_evals = 0
    for _ in range(args.steps):
        for eps in EPSRELS:
            # every rank uploads its slice of the initial region list from pinned host memory
            ws.import_regions(pin_left.numpy(), pin_len.numpy())
            r = ws.integrate(g, eps, EPSABS, subregions=True)
            h2d += 2 * pin_left.numel() * 8 * world
            tr = ws.trace()
            d2h += 64 + 256 * (len(tr) + sum(t["heuristic_probes"] for t in tr))
            e2e_evals += r.neval
    barrier()
    e2e_wall = parallel.max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_evals / e2e_wall

    # ---- trajectory parity of the timed calls against the reference's goldens (global quantities: any rank count)
    parity = [check_against_golden(FAMILY, NDIM, eps, r, tr) for eps, r, tr in collected[:len(EPSRELS)]]
    if args.mem_mode != "parity":
        parity = []
    parity_ok = bool(parity) and all(p["ok"] for p in parity)
    ws.close()

    # ---- collective side legs (every rank takes part, rank 0 reports)
    cores = os.cpu_count() or 1
    side_configs = ladder = None
    if not args.no_side_legs:
        side_configs = run_side_configs(pg, parallel, local_rank, world)
        ladder = run_native_ladder(pg, parallel, local_rank, world, args.ladder_budget_s)

    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return

    fp64_peak = pg.fp64_peak_tflops(local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    first = collected[:len(EPSRELS)]
    time_to_epsrel = []
    per_call_ms = {}
    for eps, r, tr in collected:
        per_call_ms.setdefault(eps, []).append(sum(t["eval_ms"] + t["classify_ms"] + t["heuristic_ms"] + t["split_ms"] for t in tr))
    for eps, r, tr in first:
        time_to_epsrel.append({"epsrel": eps, "status": r.status, "device_ms_kernels": min(per_call_ms[eps]),
                               "estimate": r.estimate, "errorest": r.errorest, "nregions": r.nregions,
                               "iters": r.iters, "neval": r.neval,
                               "max_regions_in_flight": max(t["nregions_global"] for t in tr)})
    peak_regions = max(t["nregions_global"] for _e, _r, tr in first for t in tr)
    out = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "F_1_8D (8-D Genz oscillatory cos(sum k*x_k)), unit cube, epsabs 1e-40, epsrel "
                        "1e-3,1e-4,...,1e-8 per step; memory model " + args.mem_mode,
            "regions_per_gpu": "contiguous slices of the region list, rebalanced when max/mean > 1.10",
            "cache": f"no L2 flush: every call restarts from 2^8 regions and the list grows to {peak_regions} "
                     f"regions ({peak_regions * 16 * NDIM / 1e6:.0f} MB of bounds, larger than the 126 MB L2)",
            "evals_per_step": evals_per_step,
        },
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d // args.steps,
                "d2h_bytes_per_step": d2h // args.steps, "wall_ms_per_step": 1e3 * e2e_wall / args.steps,
                "path": "gpuintegration_b200.Workspace.integrate (ctypes -> C ABI pgn_integrate), initial region "
                        "list uploaded from pinned host memory, integration_result + per-iteration statistics "
                        "read back (every rank uploads its slice; bytes are summed over the ranks)"},
        "gpu_launches": launches,
        "parity_ok": parity_ok if parity else None,
        "parity": {"against": "tests/golden/trajectories (unmodified reference PAGANI on a B200, oracle/_ref/ref_pagani)",
                   "checked": "per iteration: global regions, surviving regions, heuristic / budget-overflow decisions, "
                              "iteration estimate <= 1e-12; summary: status, nregions, iters, estimate <= 1e-12",
                   "cases": parity},
        "wall_ms_per_step_timed_region": 1e3 * wall / args.steps,
        "roofline": {"bound": "fp64", "kernel": "pagani_b200::detail::evaluate_kernel_bp<Oscillatory<8>>",
                     "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                     "traffic": None,
                     "note": "ALGORITHMIC flops (SURVEY 8d: 107 per F_1_8D evaluation x 1105 evaluations per region) "
                             "over the CUDA-event time of the evaluation launches of rank 0; peak = FP64 FMA "
                             "micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure); "
                             "the kernel is FP64/issue bound (arithmetic intensity ~800 flop/B; an FP64 instruction holds the issue port for two cycles, tools/evalbench/fp64_lat.cu), DRAM traffic is negligible",
                     "launches": n_eval_launches, "avg_launch_ms": eval_ms / max(n_eval_launches, 1),
                     "kernel_evals_per_s": local_evals / (eval_ms * 1e-3),
                     "share_of_device_time": eval_ms * 1e-3 / dev_s},
        "stage_ms_per_step": {"evaluate": eval_ms / args.steps, **{k[:-3]: v / args.steps for k, v in other_ms.items()}},
        "time_to_epsrel": time_to_epsrel,
        "clocks": clocks,
        "hbm_peak_gbs": peaks.get("hbm_gbs"),
    }

    if side_configs is not None:
        add_side_baselines(side_configs, cores)
        out["configs"] = side_configs
        out["native_ladder"] = ladder
    ncu = load_ncu_summary()
    if ncu:
        # executed (not contract) FP64 work and the pipe utilisation, from the committed ncu capture of this kernel
        out["roofline"]["fp64_pipe_pct"] = ncu.get("fp64_pipe_pct")
        out["roofline"]["fp64_flops_executed_per_eval"] = ncu.get("fp64_flops_executed_per_eval")
        if ncu.get("fp64_flops_executed_per_eval"):
            out["roofline"]["frac_executed"] = (local_evals * ncu["fp64_flops_executed_per_eval"] / (eval_ms * 1e-3) / 1e12
                                                / fp64_peak)
        if ncu.get("dram_bytes_per_region"):
            out["roofline"]["traffic"] = ncu["dram_bytes_per_region"] * (local_evals / feval) / max(n_eval_launches, 1)
            out["roofline"]["traffic_note"] = (
                f"dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture ({ncu.get('source')}: "
                f"{ncu.get('regions')} regions, {ncu['dram_bytes_per_region']:.1f} B per region for 16*ndim+20 = "
                f"{16 * NDIM + 20} algorithmic bytes) x regions of the average launch of this run")
        out["roofline"]["ncu_source"] = ncu.get("source")
        # The bound this kernel actually runs against: every warp instruction takes one issue cycle of
        # its scheduler and an FP64 instruction a second one; instruction mix from the same ncu capture, SM clock
        # from the samples of this run.
        try:
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            mhz = (clocks or {}).get("sm_mhz")
            wi, fi = ncu.get("warp_instructions_per_evaluation"), ncu.get("fp64_instructions_per_evaluation")
            if sms and mhz and wi and fi:
                bound = sms * 4 * mhz * 1e6 * 32.0 / (wi + fi)
                out["roofline"]["issue_bound"] = {
                    "issue_cycles_per_32_evals": wi + fi, "evals_per_s": bound,
                    "frac": out["roofline"]["kernel_evals_per_s"] / bound,
                    "note": f"{sms} SMs x 4 schedulers x {mhz:g} MHz x 32 evaluations / ({wi:.1f} warp instructions + "
                            f"{fi:.1f} second cycles of FP64 instructions per 32 evaluations)"}
        except Exception:  # a reporting extra must never cost the bench line
            pass
    if not args.no_side_legs and world == 1:
        # HBM-bound stages at the largest iteration: classify 29 B/region, fused compact+split (48 ndim + 36) B
        # per surviving region, measured peak from MEASURED_PEAKS.json (driver-written copy bandwidth)
        try:
            hbm = peaks.get("hbm_gbs") or 6650.0
            its = [t for _e, _r, tr in collected for t in tr]

            def stage_roofline(t):
                a = t["nactive"] / max(t["nregions"], 1)
                cls = 29.0 * t["nregions"] / (t["classify_ms"] * 1e-3) / 1e9 if t["classify_ms"] else None
                spl = (48 * NDIM + 36) * t["nactive"] / (t["split_ms"] * 1e-3) / 1e9 if t["split_ms"] else None
                return {"regions": t["nregions"], "active_fraction": a,
                        "classify": {"achieved": cls, "frac": cls / hbm if cls else None},
                        "compact_split": {"achieved": spl, "frac": spl / hbm if spl else None}}

            big = max(its, key=lambda t: t["nregions"])
            # compaction reads whole 32-byte sectors: with about half the regions surviving (the largest
            # iteration) most of the dropped regions' bytes are fetched anyway, so the ALGORITHMIC figure
            # (survivors only) under-states the traffic; the largest iteration in which every region
            # survives shows the kernel against the roofline without that effect
            full = [t for t in its if t["nactive"] == t["nregions"] and t["split_ms"]]
            out["roofline_hbm"] = {"bound": "hbm", "unit": "GB/s", "peak": hbm,
                                   "peak_source": "MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "fallback",
                                   "bytes_per_region": {"classify": 29, "compact_split_per_survivor": 48 * NDIM + 36},
                                   **stage_roofline(big)}
            if full:
                out["roofline_hbm"]["largest_all_active_iteration"] = stage_roofline(max(full, key=lambda t: t["nregions"]))
        except Exception as e:  # noqa: BLE001
            out["roofline_hbm"] = {"error": repr(e)}
        # CPU baseline: Cuhre (restated) on the host cores, bounded sample
        try:
            import oracle_cpu

            c = oracle_cpu.cuhre(FAMILY, NDIM, EPSRELS[0], epsabs=1e-100, maxeval=10**11, threads=cores, time_cap_s=12.0)
            out["cpu_baseline"] = {
                "value": c["neval"] / c["seconds"], "unit": "evals/s", "cores": c["threads"], "kind": "port",
                "host_cores_available": cores,
                "sample": f"Cuhre (restated, not libcuba) on F_1_8D at epsrel 1e-3, capped at 12 s "
                          f"(status {c['status']}, reached relative error estimate {c['errorest'] / abs(c['estimate']):.2e} "
                          f"after {c['neval']} evaluations)"}
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"error": repr(e)}
        # the reference's own CUDA PAGANI on the same GPU (kernel-level bar; identical trajectory => same evals)
        try:
            import refharness

            if refharness.available():
                ref = []
                for (eps, r, _tr) in first:
                    x = reference_gpu_ms(refharness.genz_name(FAMILY, NDIM), eps, repeats=4)
                    ref.append({"epsrel": eps, "ms": x["ms"], "ms_all": x["ms_all"], "status": x["status"],
                                "nregions": x["nregions"], "same_nregions_as_ours": x["nregions"] == r.nregions})
                ref_s = sum(x["ms"] for x in ref) * 1e-3
                out["reference_gpu"] = {"what": "unmodified reference CUDA PAGANI (oracle/_ref/ref_pagani, sm_100a build) "
                                                "on the same GPU, wall ms of Workspace::integrate per epsrel: minimum of "
                                                "repeats 2..4 (the first is discarded, as new_time_and_call.cuh does)",
                                        "evals_per_s": evals_per_step / ref_s, "ms_per_step": 1e3 * ref_s, "per_epsrel": ref,
                                        "speedup_e2e": e2e_value / (evals_per_step / ref_s)}
        except Exception as e:  # noqa: BLE001
            out["reference_gpu"] = {"error": repr(e)}
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
