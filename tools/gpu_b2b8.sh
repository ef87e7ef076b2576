#!/bin/bash
tag=${1:-r02u}
mkdir -p gpurun_out
{
for pairs in "down,up" "up" "down" "none"; do
echo "pairs=$pairs"; GWEN_B2B_PAIRS=$pairs timeout 200 python tools/bench_forward.py --no-torch 2>&1 | tail -1
done
} > gpurun_out/fwd_$tag.log 2>&1; cat gpurun_out/fwd_$tag.log
