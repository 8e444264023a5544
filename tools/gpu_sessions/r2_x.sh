#!/bin/bash
# round 2, session X (1 GPU): F_1_8D evaluation kernel -- uniform weights for single-set half-chunks of mixed trips, launch bounds 4..6
mkdir -p gpurun_out
: > gpurun_out/r2x_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2x_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2x_$1 $2 3 | tr -d '\n' >> gpurun_out/r2x_evalbench.jsonl; echo "}" >> gpurun_out/r2x_evalbench.jsonl; }
for rep in 1 2; do for b in runs4 runs5 noruns4 noruns5 noruns6; do run $b 5; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2x_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
