#!/bin/bash
tag=${1:-r02z2}
mkdir -p gpurun_out
timeout 300 python tools/prof_permuted.py > gpurun_out/permsweep_$tag.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/permsweep_$tag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_agg_tiled -s 2 -c 1 -f -o gpurun_out/prof_locality_$tag \
  python tools/prof_permuted.py ncu > gpurun_out/ncu_loc_$tag.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_loc_$tag.log
