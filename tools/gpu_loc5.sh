#!/bin/bash
tag=${1:-r02z5}
mkdir -p gpurun_out
log=gpurun_out/gather_sweep_$tag.log
: > $log
for cfg in "512 2" "544 2" "576 2" "576 3" "576 4" "544 1"; do
  set -- $cfg
  echo "threads=$1" >> $log
  GWEN_TILED_THREADS=$1 GWEN_GATHER_WARPS=$2 timeout 120 python tools/prof_permuted.py quick 0 >> $log 2>&1
done
GWEN_TILED_THREADS=576 GWEN_GATHER_WARPS=2 timeout 120 python tools/prof_permuted.py quick 10 256 200 >> $log 2>&1
cat $log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_agg_tiled -s 2 -c 1 -f -o gpurun_out/prof_locality_$tag \
  python tools/prof_permuted.py ncu > gpurun_out/ncu_loc_$tag.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_loc_$tag.log
