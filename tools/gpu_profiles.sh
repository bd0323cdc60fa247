#!/bin/bash
# Round profile pass: plain bench, ncu launch list of the same command, ncu --set full of the top kernels.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prof_plain.log 2>&1; echo "plain rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc $?"
timeout 600 python tools/prof_gemm.py 256 > gpurun_out/prof_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel" -s 6 -c 2 -o gpurun_out/r01_gemm_pw1 -f python tools/prof_gemm.py 256 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc $?"
wc -l gpurun_out/r01_launches.csv
