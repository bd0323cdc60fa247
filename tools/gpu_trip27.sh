#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_v0.py tests/test_gpu_attn_tc.py 2>&1 | grep -E "exit|passed|failed|^E  " | head -20
python tools/profile_v0.py 256 > gpurun_out/profile_v0_b256.log 2>&1; echo "rc $?"; head -12 gpurun_out/profile_v0_b256.log
