#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python -c "from xtag_clip_b200._cuda_probe import wait_for_cuda; print('cuda', wait_for_cuda())"
timeout 600 python scripts/tune_sweep.py --tunes ${1:-0x800,0x100800,0x200800,0x400800,0xa00,0x4800,0xc00} --iters 4 --rounds 10 > gpurun_out/s_sweep.log 2>&1; echo "sweep rc=$?"
