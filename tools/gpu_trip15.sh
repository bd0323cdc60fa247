#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_gemm_tc.py
timeout 300 python tools/prof_gemm.py 256 > gpurun_out/prof_gemm.log 2>&1; echo "prof rc $?"; tail -16 gpurun_out/prof_gemm.log
