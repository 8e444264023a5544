#!/bin/bash
# round 2, session Z (1 GPU): F_1_8D evaluation kernel -- {runs, no runs} x {launch bound 4, 5} x {cosine parity form 0, 2}, session W best as the reference
mkdir -p gpurun_out
: > gpurun_out/r2z_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2z_evalbench.jsonl; timeout 60 tools/evalbench/_build/$1 $2 3 | tr -d '\n' >> gpurun_out/r2z_evalbench.jsonl; echo "}" >> gpurun_out/r2z_evalbench.jsonl; }
for rep in 1 2; do for b in r2w_uw r2z_runs4_o0 r2z_runs4_o2 r2z_runs5_o0 r2z_runs5_o2 r2z_noruns4_o0 r2z_noruns4_o2 r2z_noruns5_o0 r2z_noruns5_o2; do run $b 5; done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
python - <<'PY'
import json
for l in open('gpurun_out/r2z_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
