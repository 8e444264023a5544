#!/bin/bash
# round 2, session E (2 GPUs): packet-based NVLink exchanges -- multi-rank tests incl. the golden-following NCCL cases,
# then the bench step on 1 and 2 ranks (stage breakdown) and the full bench line at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2e_gpus.txt
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_multirank.py > gpurun_out/r2e_pytest_multirank.log 2>&1; echo "multirank rc=$?"; tail -8 gpurun_out/r2e_pytest_multirank.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-side-legs > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench n1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-side-legs > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "bench n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2e_bench_n2_full.json 2> gpurun_out/r2e_bench_n2_full.err; echo "bench n2 full rc=$?"
python - <<'PY'
import json
for f in ("r2e_bench_n1", "r2e_bench_n2", "r2e_bench_n2_full"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, "value %.4g" % d["value"], "ms/step %.1f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "parity_ok", d.get("parity_ok"),
          {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, "wall", round(d["wall_ms_per_step_timed_region"], 1))
    for c in d.get("configs") or []:
        print("   ", c["config"], "ms %.1f" % c["ours"]["ms"], "status", c["ours"]["status"], "parity", c["parity"]["ok"], c["parity"].get("mismatches"))
    l = d.get("native_ladder")
    if l:
        print("    ladder:", l["stopped_by"], l["tightest_epsrel_reached"], [(s["epsrel"], s["status"], round(s["ms"])) for s in l["steps"]])
PY
tail -3 gpurun_out/r2e_bench_n2.err gpurun_out/r2e_bench_n2_full.err
