#!/bin/bash
# K4 check + profile (1 GPU): parity tests of the attention kernels, timings, one ncu --set full capture.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -k "xattn" > gpurun_out/x_tests.log 2>&1; echo "tests rc=$? $(tail -n 1 gpurun_out/x_tests.log)"
timeout 200 python scripts/xattn_time.py > gpurun_out/x_time.log 2>&1; echo "time rc=$?"
XTAG_TC_TUNE=0x800 timeout 300 ncu --set full --clock-control none --import-source on -k regex:xattn -c 6 -f -o gpurun_out/prof_xattn python scripts/xattn_time.py --profile > gpurun_out/x_ncu.log 2>&1; echo "ncu rc=$? $(tail -n 2 gpurun_out/x_ncu.log)"
