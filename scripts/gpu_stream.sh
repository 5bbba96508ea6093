#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python -c "from xtag_clip_b200._cuda_probe import wait_for_cuda; print('cuda', wait_for_cuda())"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -x -k "stream or fwd_blocks or cluster" > gpurun_out/st_tests.log 2>&1; echo "kernel tests rc=$? $(tail -n 1 gpurun_out/st_tests.log)"
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu --timeout 500 > gpurun_out/st_dist.log 2>&1; echo "dist rc=$? $(tail -n 1 gpurun_out/st_dist.log)"
  N=$(nvidia-smi -L | wc -l)
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/st_bench_n$N.log 2>&1; echo "bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/st_bench_n$N.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/st_bench_n$N.log | head -1)"
fi
