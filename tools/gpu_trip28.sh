#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_attn_tc.py tests/test_gpu_v0.py 2>&1 | grep -E "exit|passed|failed|^E  "
timeout 120 python tools/prof_attn.py 256 2>&1 | tail -4
python tools/profile_v0.py 256 2>&1 | grep -E "forward|attn_bias"
