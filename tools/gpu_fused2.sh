#!/bin/bash
# Fused layer kernel (round 2: L2 prefetch, k_in up to 512): parity tests, then fused vs two-kernel timings at the cfg 3
# shapes with prefetch distance 0 / 1 / 2, then the per-layer forward table.  usage: bash tools/gpu_fused2.sh <tag>
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_neighbor.py -q --timeout 600 > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/test_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke_$tag.log
for pf in 0 1 2; do
  for shape in "8 64 1024" "8 256 512" "8 512 1024" "8 512 256"; do
    GWEN_FUSED_PREFETCH=$pf timeout 120 python tools/bench_fused.py 1158 774 $shape 2>&1 | tail -1 | sed "s/^/prefetch=$pf /"
  done
done > gpurun_out/fused_$tag.log 2>&1
cat gpurun_out/fused_$tag.log
timeout 300 python tools/bench_forward.py --layers --no-torch > gpurun_out/fwd_$tag.log 2>&1; echo "fwd rc=$?"; cat gpurun_out/fwd_$tag.log
GWEN_PAIR_FUSION=0 timeout 300 python tools/bench_forward.py --no-torch > gpurun_out/fwd_nopair_$tag.log 2>&1; cat gpurun_out/fwd_nopair_$tag.log
