#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python -c "from xtag_clip_b200._cuda_probe import wait_for_cuda; print('cuda', wait_for_cuda())"
for T in 0x800 0x2800; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 8 -c 4 -f -o gpurun_out/prof_ds_$T python scripts/tune_sweep.py --tunes $T --iters 2 > gpurun_out/d_ncu_$T.log 2>&1; echo "ncu $T rc=$? $(tail -n 1 gpurun_out/d_ncu_$T.log)"
done
