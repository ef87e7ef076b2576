#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
M=1792584
{
for cfg in "" "GWEN_B2B_HB=2" "GWEN_B2B_HB=2 GWEN_B2B_S1=2" "GWEN_B2B_HB=4 GWEN_B2B_S2=4" "GWEN_B2B_S1=8 GWEN_B2B_HB=2"; do
  echo "cfg: $cfg"
  env $cfg timeout 120 python tools/bench_b2b.py $M 64 1024 512 2>&1 | tail -1
done
for cfg in "" "GWEN_B2B_S1=2" "GWEN_B2B_S1=3"; do
  echo "cfg: $cfg"
  env $cfg timeout 120 python tools/bench_b2b.py $M 512 1024 64 2>&1 | tail -1
done
} > gpurun_out/b2b_$tag.log 2>&1
cat gpurun_out/b2b_$tag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_b2b -s 2 -c 1 -f -o gpurun_out/prof_b2b_64x1024x512_$tag \
  python tools/prof_b2b.py 64 1024 512 > gpurun_out/ncu_b2b1_$tag.log 2>&1; echo "ncu b2b 1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_b2b -s 2 -c 1 -f -o gpurun_out/prof_b2b_512x1024x64_$tag \
  python tools/prof_b2b.py 512 1024 64 > gpurun_out/ncu_b2b2_$tag.log 2>&1; echo "ncu b2b 2 rc=$?"
