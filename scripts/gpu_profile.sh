#!/bin/bash
# ncu evidence for bench.py (1 GPU): launch list + one --set full capture of the tcgen05 kernels of one step.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu"
$CMD > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/bench_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 12 -c 4 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -n 3 gpurun_out/ncu_full.log
