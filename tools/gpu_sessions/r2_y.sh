#!/bin/bash
# round 2, session Y (1 GPU): F_1_8D / F_1_5D evaluation kernel -- warp-uniform warp index, per-lane weights, runs or not, launch bounds
mkdir -p gpurun_out
: > gpurun_out/r2y_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2y_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2y_$1 $2 3 | tr -d '\n' >> gpurun_out/r2y_evalbench.jsonl; echo "}" >> gpurun_out/r2y_evalbench.jsonl; }
for rep in 1 2; do for b in runs4 runs5 noruns4 noruns5 noruns6; do run $b 5; done; for b in d5_noruns4 d5_noruns5 d5_runs4; do run $b 12; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2y_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
