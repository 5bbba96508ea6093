#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -x -k "cluster or tune_bits" > gpurun_out/c_tests.log 2>&1; echo "tests rc=$? $(tail -n 1 gpurun_out/c_tests.log)"
timeout 300 python scripts/tune_sweep.py --tunes 0x800,0x4800,0x8800,0x800,0x4800,0x8800,0x800,0x4800,0x8800 --iters 5 > gpurun_out/c_sweep.log 2>&1; echo "sweep rc=$?"
