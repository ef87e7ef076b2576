#!/bin/bash
tag=${1:-r03l}
mkdir -p gpurun_out
GWEN_TC3_EPI_GROUPS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches_$tag.csv \
  python tools/prof_train_step.py > gpurun_out/ncu_train_$tag.log 2>&1; echo "ncu rc=$?"
