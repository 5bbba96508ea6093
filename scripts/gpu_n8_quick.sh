#!/bin/bash
# quick 8-GPU measurement: bench (pull exchange, graph) + timeline
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1 XTAG_SPIN_TIMEOUT_MS=30000
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/r2_bench_n${N}_c.log 2>&1
echo "bench rc=$? $(tail -1 gpurun_out/r2_bench_n${N}_c.log | cut -c1-330)"
XGRAPH=1 timeout 200 $TR --master-port 29582 scripts/timeline.py > gpurun_out/r2_timeline_n${N}_c.log 2>&1; echo "timeline rc=$?"
