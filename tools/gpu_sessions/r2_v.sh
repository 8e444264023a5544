#!/bin/bash
# round 2, session V (1 GPU): F_1_8D evaluation kernel -- launch-time range promise (no per-point checks), middle trips
# as run-time runs, unrolled middle loop, each at launch bounds 4 and 5
mkdir -p gpurun_out
: > gpurun_out/r2v_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2v_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2v_$1 $2 3 | tr -d '\n' >> gpurun_out/r2v_evalbench.jsonl; echo "}" >> gpurun_out/r2v_evalbench.jsonl; }
for rep in 1 2; do for b in base nochk runs both u2 u2nc nochk5 both5 runs5 u2nc5; do run $b 5; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2v_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
