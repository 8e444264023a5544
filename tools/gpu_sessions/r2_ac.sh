#!/bin/bash
# round 2, session AC (1 GPU): coordinate table row stride 11 / 12 / 16 doubles
mkdir -p gpurun_out
: > gpurun_out/r2ac_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2ac_evalbench.jsonl; timeout 120 tools/evalbench/_build/r2ac_$1 $2 3 | tr -d '\n' >> gpurun_out/r2ac_evalbench.jsonl; echo "}" >> gpurun_out/r2ac_evalbench.jsonl; }
for rep in 1 2; do for xs in 11 12 16; do run osc8_x$xs 5; run osc5_x$xs 12; run c010_x$xs 3; run disc10_x$xs 3; run cp8_x$xs 4; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2ac_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
