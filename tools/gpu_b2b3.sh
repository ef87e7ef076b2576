#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
M=1792584
{
for shape in "64 1024 512" "512 1024 64" "256 1024 256"; do
GWEN_B2B_PROF=1 timeout 120 python tools/bench_b2b.py $M $shape 2>&1 | tail -2
timeout 120 python tools/bench_b2b.py $M $shape 2>&1 | tail -1
done
timeout 120 python tools/bench_b2b.py 7170336 64 1024 512 2>&1 | tail -1
timeout 120 python tools/bench_b2b.py 7170336 512 1024 64 2>&1 | tail -1
} > gpurun_out/b2b_$tag.log 2>&1
cat gpurun_out/b2b_$tag.log
