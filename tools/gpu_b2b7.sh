#!/bin/bash
tag=${1:-r02t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -q --timeout 120 -x -k "b2b or pair or fused" > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/test_$tag.log
{
timeout 200 python tools/bench_forward.py --no-torch 2>&1 | tail -1
GWEN_B2B_FUSION=0 timeout 200 python tools/bench_forward.py --no-torch 2>&1 | tail -1
} > gpurun_out/fwd_$tag.log 2>&1; cat gpurun_out/fwd_$tag.log
