#!/bin/bash
# One GPU visit (round 2): parity tests, smoke, bench (both arms), GEMM sweep with 32/64-column epilogue stores,
# per-layer forward breakdown.  usage (under gpurun): bash tools/gpu_round2.sh <tag>
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/test_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_ref_$tag.json
GWEN_TC3_EPI_COLS=32 timeout 300 python tools/sweep_linear.py > gpurun_out/sweep32_$tag.log 2>&1; echo "sweep32 rc=$?"; cat gpurun_out/sweep32_$tag.log
timeout 300 python tools/sweep_linear.py > gpurun_out/sweep64_$tag.log 2>&1; echo "sweep64 rc=$?"; cat gpurun_out/sweep64_$tag.log
timeout 300 python tools/bench_forward.py --layers --no-torch > gpurun_out/fwd_$tag.log 2>&1; echo "fwd rc=$?"; cat gpurun_out/fwd_$tag.log
