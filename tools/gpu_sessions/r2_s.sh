#!/bin/bash
# round 2, session S (1 GPU): per-lane rule weights from shared memory instead of the constant bank (first F_1_8D
# alone, then with code-word prefetch variants, then every family as shipped with the option off / on)
mkdir -p gpurun_out
: > gpurun_out/r2s_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2s_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2s_$1 $2 3 | tr -d '\n' >> gpurun_out/r2s_evalbench.jsonl; echo "}" >> gpurun_out/r2s_evalbench.jsonl; }
for lw in 0 1; do run Oscillatory8_lw$lw 5; run Oscillatory5_lw$lw 12; run CZero10_lw$lw 3; run Discontinuous10_lw$lw 3; run Gaussian8_lw$lw 4; run CornerPeak8_lw$lw 4; run ProductPeak6_lw$lw 8; done
python - <<'PY'
import json
for l in open('gpurun_out/r2s_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
