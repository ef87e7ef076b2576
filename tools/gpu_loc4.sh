#!/bin/bash
tag=${1:-r02z4}
mkdir -p gpurun_out
log=gpurun_out/gather_sweep_$tag.log
: > $log
timeout 300 python -m pytest tests/test_gpu_locality.py -x -q --timeout 300 > gpurun_out/test_loc_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/test_loc_$tag.log
for cfg in "1 1 0" "1 2 0" "1 2 8 192 128" "1 2 7 176 112" "1 2 6 128 96" "1 1 6 128 96" "1 2 10 256 200"; do
  set -- $cfg
  GWEN_GATHER_MODE=$1 GWEN_GATHER_WARPS=$2 timeout 120 python tools/prof_permuted.py quick $3 $4 $5 >> $log 2>&1
done
GWEN_TILED_STAGES=3 GWEN_GATHER_WARPS=2 timeout 120 python tools/prof_permuted.py quick 6 128 96 >> $log 2>&1
GWEN_TILED_STAGES=2 GWEN_GATHER_WARPS=2 timeout 120 python tools/prof_permuted.py quick 6 128 96 >> $log 2>&1
cat $log
timeout 300 python tools/prof_permuted.py > gpurun_out/permsweep_$tag.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/permsweep_$tag.log
