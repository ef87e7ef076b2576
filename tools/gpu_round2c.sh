#!/bin/bash
# Round 2, visit c: failing-test recheck + smoke, power/backoff experiment on the full forward, epilogue staging
# knobs, then the ncu evidence (launch list of bench.py, full captures of the stencil at cfg 2 and of the fused
# kernel at 512->1024).  usage: bash tools/gpu_round2c.sh <tag>
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q --timeout 600 -k "backward or any_edge or improved" > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/test_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke_$tag.log
for ns in 0 100 400; do
  echo "GWEN_WAIT_NS=$ns"; GWEN_WAIT_NS=$ns timeout 200 python tools/bench_forward.py --no-torch 2>&1 | tail -1
done > gpurun_out/waitns_$tag.log 2>&1; cat gpurun_out/waitns_$tag.log
for eb in 1 2; do
  GWEN_FUSED_EPI_BUFS=$eb timeout 120 python tools/bench_fused.py 1158 774 8 64 1024 2>&1 | tail -1 | sed "s/^/fused_epi_bufs=$eb /"
done > gpurun_out/epibufs_$tag.log 2>&1
GWEN_TC3_BUFS=2 timeout 200 python tools/sweep_linear.py 2>&1 | sed "s/^/tc3_bufs=2 /" >> gpurun_out/epibufs_$tag.log; cat gpurun_out/epibufs_$tag.log
# --- ncu evidence (each only after the same command ran clean without ncu above / in visit a)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-model-probes --e2e-steps 2 > gpurun_out/ncu1_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_stencil -s 5 -c 2 -f -o gpurun_out/prof_stencil_$tag \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-model-probes --e2e-steps 2 > gpurun_out/ncu2_$tag.log 2>&1; echo "ncu stencil rc=$?"
timeout 300 python tools/prof_fused.py 512 1024 > gpurun_out/plain_fused_$tag.log 2>&1; echo "prof_fused plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gcn_fused -s 2 -c 1 -f -o gpurun_out/prof_fused_512x1024_$tag \
  python tools/prof_fused.py 512 1024 > gpurun_out/ncu3_$tag.log 2>&1; echo "ncu fused rc=$?"
