#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
M=1792584
run() { timeout 60 env $ENVX python tools/bench_b2b.py "$@" 2>&1 | tail -1; }
{
for e in "" "GWEN_WAIT_NS=100" "GWEN_WAIT_NS=400" "GWEN_B2B_S1=2"; do
echo "env: $e"
ENVX="$e" run $M 64 1024 512
ENVX="$e" run $M 512 1024 64
done
} > gpurun_out/b2b_$tag.log 2>&1
cat gpurun_out/b2b_$tag.log
