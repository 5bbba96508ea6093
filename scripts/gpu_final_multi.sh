#!/bin/bash
# Round-end evidence on N GPUs of one box (gpurun --gpus N): parity worker, bench (driver flags), timeline.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1 XTAG_SPIN_TIMEOUT_MS=30000
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29555 tests/dist_gpu_worker.py > gpurun_out/r2f_dist$N.log 2>&1; echo "dist rc=$? $(grep total_failures gpurun_out/r2f_dist$N.log)"; grep FAIL gpurun_out/r2f_dist$N.log | head -20
timeout 300 $TR --master-port 29571 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2f_bench_n$N.log 2>&1
echo "bench rc=$? $(tail -1 gpurun_out/r2f_bench_n$N.log | cut -c1-330)"
timeout 200 $TR --master-port 29572 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref_n$N.log 2>&1
echo "ref rc=$? $(tail -1 gpurun_out/r2f_bench_ref_n$N.log | cut -c1-200)"
XGRAPH=1 timeout 200 $TR --master-port 29582 scripts/timeline.py > gpurun_out/r2f_timeline_n$N.log 2>&1; echo "timeline rc=$?"
