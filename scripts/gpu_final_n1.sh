#!/bin/bash
# Round-end evidence on ONE GPU: smoke, bench (driver flags), reference arm, ncu launch list + full capture.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python -c "from xtag_clip_b200._cuda_probe import wait_for_cuda; print('cuda', wait_for_cuda())"
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r2f_smoke.log | cut -c1-200)"
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.log 2>&1; echo "bench rc=$? $(tail -1 gpurun_out/r2f_bench_n1.log | cut -c1-200)"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.log 2>&1; echo "ref rc=$? $(tail -1 gpurun_out/r2f_bench_ref.log | cut -c1-200)"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-extras"
$CMD > gpurun_out/r2f_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r2f_plain2.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 16 -c 4 -f -o gpurun_out/r2f_prof $CMD > gpurun_out/r2f_ncu_full.log 2>&1
echo "full capture rc=$? $(tail -n 1 gpurun_out/r2f_ncu_full.log | cut -c1-200)"
timeout 200 python scripts/sustained_ab.py --tunes 0x200800,0x1200800 --secs 1.2 --rounds 2 > gpurun_out/r2f_sustained.log 2>&1; echo "sustained $(tail -1 gpurun_out/r2f_sustained.log | cut -c1-400)"
timeout 200 python scripts/xattn_time.py > gpurun_out/r2f_xattn.log 2>&1; echo "xattn rc=$?"
