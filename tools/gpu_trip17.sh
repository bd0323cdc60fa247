#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_attn_tc.py
timeout 120 python tools/prof_attn.py 256 > gpurun_out/prof_attn.log 2>&1; echo "prof rc $?"; cat gpurun_out/prof_attn.log | tail
