#!/bin/bash
# Multi-GPU visit: usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [tests]
tag=${1:-r02m}; n=${2:-2}; tests=${3:-yes}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$tests" = "yes" ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -q --timeout 600 > gpurun_out/test_multi_$tag.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/test_multi_$tag.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n${n}_$tag.json 2> gpurun_out/bench_n${n}_$tag.err; echo "bench N=$n rc=$?"
cat gpurun_out/bench_n${n}_$tag.json; tail -12 gpurun_out/bench_n${n}_$tag.err
