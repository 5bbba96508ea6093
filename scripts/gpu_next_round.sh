#!/bin/bash
# First GPU call of the next round: validates / measures everything that was prepared without a GPU.
#   1 GPU :  bash scripts/gpu_next_round.sh
#   N GPUs:  bash scripts/gpu_next_round.sh   (under gpurun --gpus N: adds the two-stream pull A/B at N ranks)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
S=gpurun_out/next_summary.txt
: > $S
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a $S; local t0=$(date +%s)
  timeout $to "$@" > gpurun_out/$name.log 2>&1; local rc=$?
  echo "rc=$rc $(( $(date +%s) - t0 ))s $(tail -n 1 gpurun_out/$name.log | cut -c1-240)" | tee -a $S; }
python -c "from xtag_clip_b200._cuda_probe import wait_for_cuda; print('cuda', wait_for_cuda())"
N=$(nvidia-smi -L | wc -l)
# gated tests of the fusion head / DQNCOSLoss on the real kernels
XTAG_EXPERIMENTAL=1 run nx_fusion 300 python -m pytest tests/test_fusion_head.py tests/test_gpu_api.py -q -m gpu --timeout 200 -k "fusion or dqn or chunks"
# de-duplicated next-tile L2 prefetch (0x8ff) and the n-slab schedule against the default, sustained state
run nx_sweep 300 python scripts/tune_sweep.py --tunes 0x800,0x8ff,0x1008ff,0x2008ff,0x100800 --iters 4 --rounds 10
run nx_all 900 python -m pytest tests -q -m gpu --timeout 600
run nx_sweep_cfg 600 python scripts/sweep.py
if [ "$N" -ge 2 ]; then
  for PS in 1 2; do
    run nx_bench_n${N}_ps$PS 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2958$PS bench.py --gpus $N --steps 20 --warmup 5 --pull-streams $PS
  done
  # push exchange (no start-of-step barrier): parity first, then the bench line
  XTAG_EXCHANGE=push run nx_dist_push 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29591 tests/dist_gpu_worker.py
  XTAG_EXCHANGE=push run nx_bench_n${N}_push 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus $N --steps 20 --warmup 5
fi
cat $S
