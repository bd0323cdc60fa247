#!/bin/bash
mkdir -p gpurun_out
python tools/profile_v0.py 256 > gpurun_out/profile_v0_b256.log 2>&1; echo "rc $?"; head -50 gpurun_out/profile_v0_b256.log
