#!/bin/bash
tag=${1:-r02}
mkdir -p gpurun_out
timeout 120 python tools/prof_bwd.py > gpurun_out/plain_bwd_$tag.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/plain_bwd_$tag.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_linear_tc3 -s 2 -c 1 -f -o gpurun_out/prof_masked_dgrad_$tag python tools/prof_bwd.py > gpurun_out/ncu_md_$tag.log 2>&1; echo "ncu masked dgrad rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_tc -s 2 -c 1 -f -o gpurun_out/prof_wgrad_db_$tag python tools/prof_bwd.py > gpurun_out/ncu_wd_$tag.log 2>&1; echo "ncu wgrad db rc=$?"
