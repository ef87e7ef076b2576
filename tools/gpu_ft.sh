#!/bin/bash
for f in 0 1 0 1; do
  echo "GWEN_FUSED_TRAIN=$f"
  GWEN_FUSED_TRAIN=$f timeout 300 python tools/bench_train.py --iters 7 2>&1 | tail -1
done
