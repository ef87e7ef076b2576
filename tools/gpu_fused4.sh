#!/bin/bash
# Fused kernel with per-group source rings and the A block ring (one-N-tile layers): parity, timings, forward.
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q --timeout 120 -x -k "fused or pair" > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/test_$tag.log
run() {
  for shape in "8 64 1024" "8 256 512" "8 512 1024" "8 512 256" "8 256 256" "8 1024 256"; do
    env "$@" timeout 120 python tools/bench_fused.py 1158 774 $shape 2>&1 | tail -1 | sed "s/^/$* /"
  done
}
{ run GWEN_FUSED_RING=0; run GWEN_FUSED_RING=1; } > gpurun_out/fused_$tag.log 2>&1
cat gpurun_out/fused_$tag.log
for r in 0 1; do echo "GWEN_FUSED_RING=$r"; GWEN_FUSED_RING=$r timeout 200 python tools/bench_forward.py --layers --no-torch 2>&1 | tail -7; done > gpurun_out/fwd_$tag.log 2>&1; cat gpurun_out/fwd_$tag.log
