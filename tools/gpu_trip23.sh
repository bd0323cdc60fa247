#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_attn_tc.py 2>&1 | grep -E "exit|passed|failed|Error|assert" | head
python tools/profile_step.py 32 xl 384 > gpurun_out/profile_step_xl.log 2>&1; echo "profile rc $?"; head -8 gpurun_out/profile_step_xl.log
