#!/bin/bash
tag=${1:-r03p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q --timeout 300 -k "masked_dgrad or relu_links or wgrad_with_bias" > gpurun_out/test_bwd_$tag.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/test_bwd_$tag.log
for f in 1 0; do
  echo "GWEN_BWD_MASK_FUSION=$f"
  GWEN_BWD_MASK_FUSION=$f timeout 300 python tools/bench_train.py --iters 5 2>&1 | tail -1
done
