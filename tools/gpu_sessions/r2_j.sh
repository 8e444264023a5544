#!/bin/bash
# round 2, session J (2 GPUs): threshold search over the tables of the concatenated list (one all-reduce per search,
# peer exchanges only for undecided probes) -- NCCL tests against the goldens at 2 ranks, bench at 1 and 2 ranks with
# the per-search exchange counts; F_1_8D evaluation kernel register-budget variants
mkdir -p gpurun_out
: > gpurun_out/r2j_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2j_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2j_$1 $2 3 | tr -d '\n' >> gpurun_out/r2j_evalbench.jsonl; echo "}" >> gpurun_out/r2j_evalbench.jsonl; }
for b in osc8_mb4 osc8_mb5 osc8_mb4_np osc8_mb5_np osc8_mb5_tailA osc8_mb5 osc8_mb4; do run $b 5; done
python - <<'PY'
import json
for l in open('gpurun_out/r2j_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_search.py tests/test_gpu_heuristic_kat.py -m gpu -q -x > gpurun_out/r2j_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest_multi.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-side-legs > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench n=1 rc=$?"
PGN_SEARCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-side-legs > gpurun_out/r2j_bench_n2_debug.json 2> gpurun_out/r2j_bench_n2_debug.err; echo "bench n=2 (debug) rc=$?"
grep "pgn device search" gpurun_out/r2j_bench_n2_debug.err | python -c "
import sys,re
p=c=x=n=0
for l in sys.stdin:
    m=re.search(r'probes=(\d+) computed=(\d+) peer exchanges=(\d+)', l)
    if m: n+=1; p+=int(m.group(1)); c+=int(m.group(2)); x+=int(m.group(3))
print('searches (both ranks, all steps)', n, 'probes', p, 'computed', c, 'exchanges', x)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "bench n=2 rc=$?"
python - <<'PY'
import json
for n in (1, 2):
    try:
        d=json.loads([l for l in open('gpurun_out/r2j_bench_n%d.json' % n) if l.startswith('{')][-1])
        print(n, 'value %.4g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'parity', d.get('parity_ok'), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'launches', d['gpu_launches'])
    except Exception as e: print(n, 'no line', e)
PY
tail -3 gpurun_out/r2j_bench_n2.err | cut -c1-300
