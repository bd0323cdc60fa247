#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py tests/test_gpu_model.py
python tools/profile_step.py 256 > gpurun_out/profile_step_b256.log 2>&1; echo "profile rc $?"; head -80 gpurun_out/profile_step_b256.log
