#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
M=1792584
{
for shape in "1000 64 256 64" "70001 320 512 192" "$M 64 1024 512" "$M 512 1024 64" "$M 256 1024 256"; do
GWEN_B2B_PROF=1 timeout 120 python tools/bench_b2b.py $shape 2>&1 | tail -2
timeout 120 python tools/bench_b2b.py $shape 2>&1 | tail -1
done
timeout 120 python tools/bench_b2b.py 7170336 64 1024 512 2>&1 | tail -1
timeout 120 python tools/bench_b2b.py 7170336 512 1024 64 2>&1 | tail -1
} > gpurun_out/b2b_$tag.log 2>&1
cat gpurun_out/b2b_$tag.log
