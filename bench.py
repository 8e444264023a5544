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
        r = oracle_cpu.cuhre(FAMILY, NDIM, EPSRELS[0], epsabs=1e-100, maxeval=10**11, threads=0, time_cap_s=cap)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mem-mode", default="parity", choices=["parity", "native"])
    ap.add_argument("--no-side-legs", action="store_true", help="skip cpu_baseline / reference-GPU legs")
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
            if world == 1:
                ws.import_regions(pin_left.numpy(), pin_len.numpy())
                r = ws.integrate(g, eps, EPSABS, subregions=True)
                h2d += 2 * pin_left.numel() * 8
            else:
                r = ws.integrate(g, eps, EPSABS)
            tr = ws.trace()
            d2h += 64 + 256 * (len(tr) + sum(t["heuristic_probes"] for t in tr))
            e2e_evals += r.neval
    barrier()
    e2e_wall = parallel.max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_evals / e2e_wall

    if rank != 0:
        ws.close()
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
                        "read back" if world == 1 else "gpuintegration_b200.Workspace.integrate on every rank"},
        "gpu_launches": launches,
        "wall_ms_per_step_timed_region": 1e3 * wall / args.steps,
        "roofline": {"bound": "fp64", "kernel": "pagani_b200::detail::evaluate_kernel_bp<Oscillatory<8>>",
                     "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                     # DRAM bytes per launch: 128.9 B per region (ncu, below) x regions of the average launch
                     "traffic": 128.9 * (local_evals / feval) / max(n_eval_launches, 1),
                     "traffic_note": "ncu --set full, 65536-region launch (profiles/r01_ncu_eval_s8_final.txt): 8.45 MB read, "
                                     "0 B written back before the kernel ends = 129 B per region for 16*ndim+20 = 148 "
                                     "algorithmic bytes (part of the bounds is still in L2 from the split kernel)",
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

            c = oracle_cpu.cuhre(FAMILY, NDIM, EPSRELS[0], epsabs=1e-100, maxeval=10**11, threads=0, time_cap_s=12.0)
            out["cpu_baseline"] = {
                "value": c["neval"] / c["seconds"], "unit": "evals/s", "cores": c["threads"], "kind": "port",
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
                    lines = refharness.integrate(refharness.genz_name(FAMILY, NDIM), eps, EPSABS, repeats=2)
                    ref.append({"epsrel": eps, "ms": lines[-1]["ms"], "status": lines[-1]["status"],
                                "nregions": lines[-1]["nregions"], "same_nregions_as_ours": lines[-1]["nregions"] == r.nregions})
                ref_s = sum(x["ms"] for x in ref) * 1e-3
                out["reference_gpu"] = {"what": "unmodified reference CUDA PAGANI (oracle/_ref/ref_pagani, sm_100a build) "
                                                "on the same GPU, wall ms of Workspace::integrate per epsrel, 2nd repeat",
                                        "evals_per_s": evals_per_step / ref_s, "ms_per_step": 1e3 * ref_s, "per_epsrel": ref,
                                        "speedup_e2e": e2e_value / (evals_per_step / ref_s)}
        except Exception as e:  # noqa: BLE001
            out["reference_gpu"] = {"error": repr(e)}
    ws.close()
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
