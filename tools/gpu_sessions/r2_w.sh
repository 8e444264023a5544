#!/bin/bash
# round 2, session W (1 GPU): F_1_8D evaluation kernel -- parity predicate, low-word-first shuffles, warp-uniform warp index
mkdir -p gpurun_out
: > gpurun_out/r2w_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2w_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2w_$1 $2 3 | tr -d '\n' >> gpurun_out/r2w_evalbench.jsonl; echo "}" >> gpurun_out/r2w_evalbench.jsonl; }
for rep in 1 2; do for b in v_both odd2 odd2lo uw uwodd2 uwodd2_5 uwnoruns uwnoruns5 uwchk; do run $b 5; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2w_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
