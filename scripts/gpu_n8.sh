#!/bin/bash
# 8-GPU shot: parity of the multi-rank paths, the headline bench line, one timeline of a captured step.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
N=$(nvidia-smi -L | wc -l)
python -c "from xtag_clip_b200._cuda_probe import wait_for_cuda; print('cuda', wait_for_cuda())"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/dist_gpu_worker.py > gpurun_out/n8_dist.log 2>&1; echo "dist rc=$? $(grep total_failures gpurun_out/n8_dist.log)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n8_bench.log 2>&1; echo "bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/n8_bench.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/n8_bench.log | head -1)"
XGRAPH=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 scripts/timeline.py > gpurun_out/n8_timeline.log 2>&1; echo "timeline rc=$?"
