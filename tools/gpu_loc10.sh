#!/bin/bash
tag=${1:-r03f}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_locality.py -x -q --timeout 300 > gpurun_out/test_loc_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/test_loc_$tag.log
GWEN_BENCH_TRI=1 timeout 300 python tools/bench_permuted.py > gpurun_out/permuted_tri_$tag.json 2> gpurun_out/permuted_tri_$tag.err; echo "tri rc=$?"; cat gpurun_out/permuted_tri_$tag.json; tail -3 gpurun_out/permuted_tri_$tag.err
