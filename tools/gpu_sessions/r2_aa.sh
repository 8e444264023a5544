#!/bin/bash
# round 2, session AA (1 GPU): every Genz family, warp-uniform warp index off / on, unit-Jacobian variant off / on for the 10-D families
mkdir -p gpurun_out
: > gpurun_out/r2aa_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2aa_evalbench.jsonl; timeout 120 tools/evalbench/_build/r2aa_$1 $2 3 | tr -d '\n' >> gpurun_out/r2aa_evalbench.jsonl; echo "}" >> gpurun_out/r2aa_evalbench.jsonl; }
for rep in 1 2; do for uw in 0 1; do
run osc8_uw$uw 5; run osc5_mb5_uw$uw 12; run osc5_mb4_uw$uw 12
for uj in 0 1; do run c010_uj${uj}_uw$uw 3; run disc10_uj${uj}_uw$uw 3; done
run gauss8_uw$uw 4; run gauss5_uw$uw 9; run pp6_uw$uw 8; run cp8_uw$uw 4
done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2aa_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
