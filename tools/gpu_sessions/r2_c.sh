#!/bin/bash
# round 2, session C: evaluation kernel with the call-free batch_fast protocol vs the round-1 kernel; is the xy_2 golden
# mismatch new?; full GPU suite with the new library; bench; ncu
mkdir -p gpurun_out
: > gpurun_out/r2c_evalbench.jsonl
run() { # binary parts
  echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2c_evalbench.jsonl
  timeout 60 tools/evalbench/_build/$1 $2 3 | tr -d '\n' >> gpurun_out/r2c_evalbench.jsonl; echo "}" >> gpurun_out/r2c_evalbench.jsonl
}
for b in tools/evalbench/_build/r2_*osc8*; do run $(basename $b) 5; done
for b in tools/evalbench/_build/r2_*Gaussian8 tools/evalbench/_build/r2_*CornerPeak8; do run $(basename $b) 4; done
for b in tools/evalbench/_build/r2_*ProductPeak6; do run $(basename $b) 8; done
for b in tools/evalbench/_build/r2_*10; do run $(basename $b) 3; done
for b in tools/evalbench/_build/r2_*Gaussian5; do run $(basename $b) 12; done
python - <<'PY'
import json
for l in open('gpurun_out/r2c_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
# the 2-D x*y golden with the ROUND-1 library (then the new one is put back)
cp gpuintegration_b200/libpagani_b200.so /tmp/new_lib.so
cp tools/evalbench/_build/oldlib/libpagani_b200.so gpuintegration_b200/libpagani_b200.so
timeout 120 python -m pytest -q -m gpu "tests/test_gpu_parity.py::test_trajectory_matches_golden[xy_2_1e-06]" "tests/test_gpu_parity.py::test_per_region_arrays_match_golden[xy_2_1e-06]" "tests/test_gpu_parity.py::test_trajectory_matches_golden[const_2_0.001]" > gpurun_out/r2c_xy2_oldlib.log 2>&1; echo "xy2 with old lib rc=$?"; tail -4 gpurun_out/r2c_xy2_oldlib.log
cp /tmp/new_lib.so gpuintegration_b200/libpagani_b200.so
timeout 1200 python -m pytest -q -m gpu tests > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2c_pytest.log
timeout 200 python bench.py --steps 2 --warmup 1 --no-side-legs > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2c_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2c_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
