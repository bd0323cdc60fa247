#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py
timeout 300 python tools/prof_ops.py 256 > gpurun_out/prof_ops.log 2>&1; echo "prof rc $?"; cat gpurun_out/prof_ops.log
