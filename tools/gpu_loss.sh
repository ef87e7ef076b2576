#!/bin/bash
tag=${1:-r03s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "masked_l1 or train_step" > gpurun_out/test_loss_$tag.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/test_loss_$tag.log
timeout 300 python tools/bench_train.py --iters 5 2>&1 | tail -1
