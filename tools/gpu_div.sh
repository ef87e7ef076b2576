#!/bin/bash
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q --timeout 120 -x -k "fused or pair or linear or b2b" > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/test_$tag.log
for shape in "8 512 1024" "8 64 1024" "8 256 512" "8 512 256"; do
  timeout 120 python tools/bench_fused.py 1158 774 $shape 2>&1 | tail -1
done > gpurun_out/fused_$tag.log 2>&1
cat gpurun_out/fused_$tag.log
timeout 200 python tools/sweep_linear.py > gpurun_out/sweep_$tag.log 2>&1; cat gpurun_out/sweep_$tag.log
timeout 200 python tools/bench_forward.py --no-torch > gpurun_out/fwd_$tag.log 2>&1; cat gpurun_out/fwd_$tag.log
