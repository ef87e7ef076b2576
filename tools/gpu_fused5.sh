#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
GWEN_FUSED_RING=2 timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q --timeout 120 -x -k "fused or pair" > gpurun_out/test_$tag.log 2>&1; echo "pytest(ring=2) rc=$?"; tail -6 gpurun_out/test_$tag.log
run() {
  for shape in "8 256 512" "8 512 256" "8 128 512"; do
    env "$@" timeout 120 python tools/bench_fused.py 1158 774 $shape 2>&1 | tail -1 | sed "s/^/$* /"
  done
}
{ run GWEN_FUSED_RING=1; run GWEN_FUSED_RING=2; run GWEN_FUSED_RING=2 GWEN_FUSED_SG=1; run GWEN_FUSED_RING=1 GWEN_FUSED_SG=1; } > gpurun_out/fused_$tag.log 2>&1
cat gpurun_out/fused_$tag.log
for cfg in "GWEN_FUSED_RING=1" "GWEN_FUSED_RING=2" "GWEN_FUSED_RING=2 GWEN_FUSED_SG=1" "GWEN_FUSED_RING=1 GWEN_FUSED_MIN_K=128" "GWEN_FUSED_RING=2 GWEN_FUSED_MIN_K=128"; do
  echo "$cfg"; env $cfg timeout 200 python tools/bench_forward.py --no-torch 2>&1 | tail -1
done > gpurun_out/fwd_$tag.log 2>&1; cat gpurun_out/fwd_$tag.log
