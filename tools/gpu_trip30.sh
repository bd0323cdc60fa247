#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_v0.py 2>&1 | grep -E "exit|passed|failed|^E  " | head
python tools/profile_v0.py 256 2>&1 | head -14
