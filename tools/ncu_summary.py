"""Compact summary of an Nsight Compute report (ncu -i rep --page raw --csv): the metrics the roofline
discussion in DESIGN.md uses.  usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
def summarize(reps):
  for rep in reps:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print(f"== {rep}: {d.get('Kernel Name', '?')[:110]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>18s} {u[k]}")
        stalls = {k.split('stalled_')[1].split('_per')[0]: float(d[k]) for k in hdr
                  if 'issue_stalled' in k and 'per_issue_active' in k and d[k]}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:8]
        print("  warp stall reasons (avg warps stalled per issue slot): " + ", ".join(f"{k}={v:.2f}" for k, v in top))


def eval_kernel_json(rep, regions, feval, out_path):
    """Figures bench.py quotes next to the evaluation kernel's roofline, from one `ncu --set full --import-source on`
    capture of a launch over `regions` regions: FP64 pipe utilisation, executed warp instructions per evaluation by
    opcode (source page), executed FP64 flops per evaluation (DFMA = 2, DADD/DMUL/DSETP = 1), DRAM bytes per region."""
    import collections
    import json
    import re

    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
    hdr = srows[hi]
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    per = regions * feval / 32.0  # warp instructions per "evaluation" = per 32 points
    ops = collections.Counter()
    for r in srows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
        ops[op] += int(r[iex])
    flops = (2 * ops["DFMA"] + ops["DADD"] + ops["DMUL"] + ops["DSETP"] + ops["DMNMX"]) / per
    fp64 = sum(ops[k] for k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")) / per
    f = lambda k: float(d[k].replace(",", "")) if d.get(k) else None  # noqa: E731
    doc = {
        "source": rep, "kernel": d.get("Kernel Name"), "regions": regions, "feval": feval,
        "duration_us": f("gpu__time_duration.sum"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lsu_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "warps_active_per_sm": f("sm__warps_active.avg.per_cycle_active"),
        "warp_instructions_per_evaluation": sum(ops.values()) / per,
        "fp64_instructions_per_evaluation": fp64,
        "fp64_flops_executed_per_eval": flops,
        "opcode_histogram_per_evaluation": {k: round(v / per, 3) for k, v in ops.most_common(24)},
        "dram_bytes_read": f("dram__bytes_read.sum"), "dram_bytes_write": f("dram__bytes_write.sum"),
    }
    if doc["dram_bytes_read"] is not None:
        unit = dict(zip(rows[0], rows[1])).get("dram__bytes_read.sum", "byte")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        doc["dram_bytes_read"] *= scale
        wunit = dict(zip(rows[0], rows[1])).get("dram__bytes_write.sum", "byte")
        doc["dram_bytes_write"] *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(wunit, 1.0)
        doc["dram_bytes_per_region"] = (doc["dram_bytes_read"] + doc["dram_bytes_write"]) / regions
    json.dump(doc, open(out_path, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    # python tools/ncu_summary.py report.ncu-rep [...]                      text summary
    # python tools/ncu_summary.py --json out.json report.ncu-rep REGIONS FEVAL   evaluation-kernel figures for bench.py
    if len(sys.argv) > 1 and sys.argv[1] == "--json":
        eval_kernel_json(sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), sys.argv[2])
    else:
        summarize(sys.argv[1:])
