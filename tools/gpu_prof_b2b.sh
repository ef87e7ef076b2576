#!/bin/bash
tag=${1:-r02w}
mkdir -p gpurun_out
timeout 100 python tools/prof_b2b.py 64 1024 512 > gpurun_out/plain_b2b_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_b2b -s 2 -c 1 -f -o gpurun_out/prof_b2b_64x1024x512_$tag \
  python tools/prof_b2b.py 64 1024 512 > gpurun_out/ncu_b2b1_$tag.log 2>&1; echo "ncu b2b 1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_b2b -s 2 -c 1 -f -o gpurun_out/prof_b2b_512x1024x64_$tag \
  python tools/prof_b2b.py 512 1024 64 > gpurun_out/ncu_b2b2_$tag.log 2>&1; echo "ncu b2b 2 rc=$?"
