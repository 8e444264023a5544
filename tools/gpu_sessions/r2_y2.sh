#!/bin/bash
# round 2, session Y2 (1 GPU): session W's and session Y's binaries on the same box (box-to-box differences are ~1 %)
mkdir -p gpurun_out
: > gpurun_out/r2y2_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2y2_evalbench.jsonl; timeout 60 tools/evalbench/_build/$1 $2 3 | tr -d '\n' >> gpurun_out/r2y2_evalbench.jsonl; echo "}" >> gpurun_out/r2y2_evalbench.jsonl; }
for rep in 1 2 3; do for b in r2w_v_both r2w_uw r2w_uwnoruns5 r2y_runs4 r2y_noruns5 r2x_runs5 r2x_noruns5; do run $b 5; done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
python - <<'PY'
import json
for l in open('gpurun_out/r2y2_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
