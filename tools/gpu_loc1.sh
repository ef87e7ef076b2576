#!/bin/bash
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_locality.py -x -q --timeout 300 > gpurun_out/test_loc_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/test_loc_$tag.log
for args in "0" "8" "10" "8 192 128" "7 176 112"; do
  echo "== bench_permuted $args"
  timeout 300 python tools/bench_permuted.py $args > gpurun_out/perm_${tag}_$(echo $args | tr ' ' '_').json 2> gpurun_out/perm_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/perm_${tag}_$(echo $args | tr ' ' '_').json"))
    print(d["locality_plan"]); print({k: d["fp32"][k] for k in d["fp32"]}); print({k: d["bf16"][k] for k in d["bf16"]})
except Exception as e:
    print("no result", e); print(open("gpurun_out/perm_$tag.err").read()[-2000:])
PY
done
