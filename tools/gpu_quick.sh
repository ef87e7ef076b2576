#!/bin/bash
tag=${1:-r02q}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q --timeout 120 -k "$*" > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/test_$tag.log
