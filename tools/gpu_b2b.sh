#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
for shape in "1000 64 128 64" "4096 128 256 128" "70000 512 1024 64" "70001 64 1024 512" "7170336 512 1024 64" "7170336 64 1024 512" "7170336 256 1024 256"; do
  timeout 120 python tools/bench_b2b.py $shape 2>&1 | tail -2
done > gpurun_out/b2b_$tag.log 2>&1
cat gpurun_out/b2b_$tag.log
