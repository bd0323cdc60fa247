#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py tests/test_gpu_model.py
python tools/prof_gemm.py 256 > gpurun_out/prof_gemm.log 2>&1 && cat gpurun_out/prof_gemm.log && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 4 -c 2 -o gpurun_out/prof_gemm python tools/prof_gemm.py 256 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc $?"
