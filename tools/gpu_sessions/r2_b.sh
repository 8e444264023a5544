#!/bin/bash
# round 2, session B: software-pipelined evaluation kernel vs the round-1 kernel (same sums = same per-region values),
# fixed issue-slot microbenchmark, parity tests with the new library, ncu of the new kernel
mkdir -p gpurun_out
timeout 120 tools/evalbench/_build/issue_mix > gpurun_out/r2b_issue_mix.jsonl 2>&1; echo "issue_mix rc=$?"
: > gpurun_out/r2b_evalbench.jsonl
run() { # binary parts
  echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2b_evalbench.jsonl
  timeout 60 tools/evalbench/_build/$1 $2 3 | tr -d '\n' >> gpurun_out/r2b_evalbench.jsonl; echo "}" >> gpurun_out/r2b_evalbench.jsonl
}
for b in tools/evalbench/_build/r2_*osc8*; do run $(basename $b) 5; done
for b in tools/evalbench/_build/r2_*Gaussian8 tools/evalbench/_build/r2_*CornerPeak8; do run $(basename $b) 4; done
for b in tools/evalbench/_build/r2_*ProductPeak6; do run $(basename $b) 8; done
for b in tools/evalbench/_build/r2_*10; do run $(basename $b) 3; done
for b in tools/evalbench/_build/r2_*Gaussian5; do run $(basename $b) 12; done
python - <<'PY'
import json
for l in open('gpurun_out/r2b_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_gpu_integrate.py tests/test_gpu_stages.py tests/test_gpu_dropin.py > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2b_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
timeout 120 python bench.py --steps 2 --warmup 1 --no-side-legs > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2b_bench.json
