#!/bin/bash
tag=${1:-r02z8}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_locality.py -x -q --timeout 300 > gpurun_out/test_loc_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/test_loc_$tag.log
timeout 300 python tools/bench_permuted.py > gpurun_out/permuted_$tag.json 2> gpurun_out/permuted_$tag.err; echo "permuted rc=$?"; cat gpurun_out/permuted_$tag.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_agg_tiled -s 2 -c 1 -f -o gpurun_out/prof_locality_$tag \
  python tools/prof_permuted.py ncu > gpurun_out/ncu_loc_$tag.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_loc_$tag.log
timeout 300 python tools/bench_forward.py --no-torch --layers > gpurun_out/fwd_$tag.log 2>&1; echo "fwd rc=$?"; tail -12 gpurun_out/fwd_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fwd_launches_$tag.csv \
  python tools/bench_forward.py --no-torch > gpurun_out/ncu_fwd_$tag.log 2>&1; echo "ncu fwd rc=$?"
