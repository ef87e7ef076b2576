#!/bin/bash
tag=${1:-r03b}
mkdir -p gpurun_out
timeout 200 python tools/prof_train_step.py > gpurun_out/train_plain_$tag.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/train_plain_$tag.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches_$tag.csv \
  python tools/prof_train_step.py > gpurun_out/ncu_train_$tag.log 2>&1; echo "ncu rc=$?"
