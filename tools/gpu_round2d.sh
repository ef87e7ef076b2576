#!/bin/bash
# Round 2, visit d: whole GPU suite on the final tree, smoke, both bench arms, masked-mesh numbers.
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/test_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/smoke_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; cut -c1-300 gpurun_out/bench_ref_$tag.json
timeout 300 python tools/bench_masked.py 0.1 > gpurun_out/masked_$tag.json 2>&1; echo "masked rc=$?"; cat gpurun_out/masked_$tag.json
