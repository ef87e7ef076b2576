#!/bin/bash
tag=${1:-r02z6}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/test_$tag.log
log=gpurun_out/gather_sweep_$tag.log
: > $log
for cfg in "1 2 0" "1 1 0" "1 3 0" "1 2 10 256 200" "1 2 7 176 112" "1 2 6 128 96"; do
  set -- $cfg
  GWEN_GATHER_MODE=$1 GWEN_GATHER_WARPS=$2 timeout 120 python tools/prof_permuted.py quick $3 $4 $5 >> $log 2>&1
done
cat $log
timeout 300 python tools/prof_permuted.py > gpurun_out/permsweep_$tag.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/permsweep_$tag.log
timeout 300 python tools/bench_masked.py > gpurun_out/masked_$tag.json 2>&1; echo "masked rc=$?"; grep -A3 "tiled\|auto" gpurun_out/masked_$tag.json | grep -v "^--" | head -30
