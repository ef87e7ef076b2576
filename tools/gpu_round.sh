#!/bin/bash
# One GPU visit: parity tests, bench, ncu launch list, one full ncu capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/test_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-model-probes --e2e-steps 2 > gpurun_out/ncu1_$tag.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_grid_stencil -s 5 -c 2 -f -o gpurun_out/prof_stencil_$tag \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-model-probes --e2e-steps 2 > gpurun_out/ncu2_$tag.log 2>&1; echo "ncu full rc=$?"
