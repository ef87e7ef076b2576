#!/bin/bash
tag=${1:-r02z3}
mkdir -p gpurun_out
log=gpurun_out/gather_sweep_$tag.log
: > $log
timeout 300 python -m pytest tests/test_gpu_locality.py -x -q --timeout 300 > gpurun_out/test_loc_$tag.log 2>&1; echo "pytest(default mode) rc=$?"; tail -3 gpurun_out/test_loc_$tag.log
GWEN_GATHER_MODE=2 timeout 300 python -m pytest tests/test_gpu_locality.py -x -q --timeout 300 > gpurun_out/test_loc2_$tag.log 2>&1; echo "pytest(bulk mode) rc=$?"; tail -3 gpurun_out/test_loc2_$tag.log
for cfg in "1 1" "1 2" "1 3" "1 4" "2 1"; do
  set -- $cfg
  GWEN_GATHER_MODE=$1 GWEN_GATHER_WARPS=$2 timeout 120 python tools/prof_permuted.py quick 0 >> $log 2>&1
done
for cfg in "2 1 8 192 128" "1 2 8 192 128" "2 1 7 176 112" "1 2 7 176 112" "2 1 6 128 96" "1 2 6 128 96" "1 4 6 128 96"; do
  set -- $cfg
  GWEN_GATHER_MODE=$1 GWEN_GATHER_WARPS=$2 timeout 120 python tools/prof_permuted.py quick $3 $4 $5 >> $log 2>&1
done
cat $log
