#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_gemm.py 256 > gpurun_out/prof_gemm.log 2>&1; echo "prof rc $?"; tail -16 gpurun_out/prof_gemm.log
