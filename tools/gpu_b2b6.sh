#!/bin/bash
# careful: stop at the first case that does not finish
tag=${1:-r02s}
mkdir -p gpurun_out
M=1792584
run() { timeout 60 env $ENVX python tools/bench_b2b.py "$@" 2>&1 | tail -${TL:-1}; return ${PIPESTATUS[0]}; }
{
run 256 512 1024 64 || { echo HANG; exit 1; }
run 5000 512 1024 64 || { echo HANG; exit 1; }
ENVX="GWEN_B2B_HB=2" run 5000 64 1024 512 || { echo HANG; exit 1; }
run 70001 320 512 192 || { echo HANG; exit 1; }
run 70001 256 1024 256 || { echo HANG; exit 1; }
run 70001 64 1024 512 || { echo HANG; exit 1; }
for shape in "$M 64 1024 512" "$M 512 1024 64" "$M 256 1024 256"; do
ENVX="GWEN_B2B_PROF=1" TL=2 run $shape || { echo HANG; exit 1; }
run $shape || { echo HANG; exit 1; }
done
run 7170336 64 1024 512 || { echo HANG; exit 1; }
run 7170336 512 1024 64 || { echo HANG; exit 1; }
} > gpurun_out/b2b_$tag.log 2>&1
cat gpurun_out/b2b_$tag.log
