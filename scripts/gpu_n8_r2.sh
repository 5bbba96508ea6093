#!/bin/bash
# 8-GPU validation + measurement of round 2 (one gpurun --gpus 8 call)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1 XTAG_SPIN_TIMEOUT_MS=30000
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29555 tests/dist_gpu_worker.py > gpurun_out/r2_dist$N.log 2>&1; echo "dist rc=$? $(grep total_failures gpurun_out/r2_dist$N.log)"; grep FAIL gpurun_out/r2_dist$N.log | head -20
for EX in pull push; do
  timeout 200 $TR --master-port 2957$((RANDOM % 10)) bench.py --gpus $N --steps 40 --warmup 10 --exchange $EX > gpurun_out/r2_bench_n${N}_$EX.log 2>&1
  echo "bench $EX rc=$? $(tail -1 gpurun_out/r2_bench_n${N}_$EX.log | cut -c1-330)"
done
XGRAPH=1 XEXCH=push timeout 200 $TR --master-port 29581 scripts/timeline.py > gpurun_out/r2_timeline_n${N}_push.log 2>&1; echo "timeline rc=$?"
XGRAPH=1 XEXCH=pull timeout 200 $TR --master-port 29582 scripts/timeline.py > gpurun_out/r2_timeline_n${N}_pull.log 2>&1; echo "timeline rc=$?"
