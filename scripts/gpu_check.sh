#!/bin/bash
# Runs the GPU test groups in separate processes (a faulting kernel poisons only its own group), then smoke + bench.
# Usage under gpurun:  bash scripts/gpu_check.sh [quick]
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run t_simple 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -k "l2norm or lse_combine or asl or xattn"
run t_gemm 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -k "tc_gemm"
run t_clip_simt 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -k "clip_fwd_bwd and not 2-dtype2"
run t_clip_tc 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -k "(clip_fwd_bwd and 2-dtype2) or wide_dynamic"
run t_api 1200 python -m pytest tests/test_gpu_api.py -q -m gpu --timeout 600
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
if [ "$1" != "quick" ]; then
  run bench 900 python bench.py --steps 10 --warmup 3
  run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
fi
cat gpurun_out/summary.txt
