#!/bin/bash
# round 2, session P (1 GPU): ncu source-level profile of the threshold-search kernel (where do the 0.45 us per
# memo-table probe go?)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_search -s 1 -c 2 -o gpurun_out/r2p_search python tools/ncu_stages.py 1 8 1e-4 > gpurun_out/r2p_ncu_search.log 2>&1; echo "ncu search rc=$?"
tail -5 gpurun_out/r2p_ncu_search.log
ls -la gpurun_out | grep r2p
