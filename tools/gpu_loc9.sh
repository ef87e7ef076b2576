#!/bin/bash
tag=${1:-r03d}
mkdir -p gpurun_out
log=gpurun_out/gather_sweep_$tag.log
: > $log
GWEN_GATHER_MODE=3 timeout 300 python -m pytest tests/test_gpu_locality.py -x -q --timeout 300 > gpurun_out/test_loc3_$tag.log 2>&1; echo "pytest(mode 3) rc=$?"; tail -3 gpurun_out/test_loc3_$tag.log
GWEN_GATHER_MODE=1 GWEN_GATHER_WARPS=2 timeout 120 python tools/prof_permuted.py quick 0 >> $log 2>&1
for cfg in "64 2" "128 2" "192 2" "256 2" "128 3" "192 3"; do
  set -- $cfg
  echo "frac=$1/256" >> $log
  GWEN_GATHER_MODE=3 GWEN_GATHER4_FRAC=$1 GWEN_GATHER_WARPS=$2 timeout 120 python tools/prof_permuted.py quick 0 >> $log 2>&1
done
cat $log
