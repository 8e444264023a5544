#!/bin/bash
# round 2, session U (1 GPU): F_1_8D evaluation kernel -- magic-number rounding instead of F2I/I2F in the cosine,
# unit-Jacobian kernel variant (no f * 1.0)
mkdir -p gpurun_out
: > gpurun_out/r2u_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2u_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2u_$1 $2 3 | tr -d '\n' >> gpurun_out/r2u_evalbench.jsonl; echo "}" >> gpurun_out/r2u_evalbench.jsonl; }
for b in osc8_base osc8_magic osc8_unit osc8_both osc8_base osc8_magic osc8_unit osc8_both; do run $b 5; done
python - <<'PY'
import json
for l in open('gpurun_out/r2u_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
