#!/bin/bash
# Second-session GPU check (1 GPU): new kernels' parity tests first, then A/B timings of the tuning bits, then the
# whole GPU suite and a bench line.  Every step is bounded by its own timeout; logs go to gpurun_out/.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
S=gpurun_out/summary_b.txt
: > $S
run() { # name, timeout, cmd...
  local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a $S
  local t0=$(date +%s)
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 ))s  $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a $S
}
run b_new_tests 420 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -k "tune_bits or one_exp or fwd_blocks or xattn"
run b_tune_sweep 240 python scripts/tune_sweep.py
run b_xattn_time 180 python scripts/xattn_time.py
run b_all_tests 900 python -m pytest tests -q -m gpu --timeout 600
run b_smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
run b_bench 400 python bench.py --steps 10 --warmup 3
cat $S
