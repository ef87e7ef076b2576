#!/bin/bash
# multi-GPU e2e diagnostic: bench headline + e2e only, with and without NUMA placement of the pinned buffers
tag=${1:-r02e}; n=${2:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$tag.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" > gpurun_out/lscpu_$tag.txt 2>&1
for mode in numa nonuma; do
  if [ $mode = nonuma ]; then export GWEN_NO_NUMA=1; else unset GWEN_NO_NUMA; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $n --steps 20 --warmup 5 --no-model-probes > gpurun_out/bench_e2e_${mode}_n${n}_$tag.json 2> gpurun_out/bench_e2e_${mode}_$tag.err; echo "bench $mode rc=$?"
  python - <<PY
import json
j=json.load(open("gpurun_out/bench_e2e_${mode}_n${n}_$tag.json"))
print("$mode", j["e2e"]["value"], j["e2e"]["gbs_per_direction_per_gpu"], [(r["rank"], r["gpu_numa_node"], r["node_cpus"], r["allowed_cpus"], round(r["e2e_ms_per_step"],2)) for r in j["e2e"]["per_rank"]])
PY
done
cat gpurun_out/lscpu_$tag.txt; head -14 gpurun_out/topo_$tag.txt
