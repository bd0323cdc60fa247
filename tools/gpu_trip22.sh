#!/bin/bash
mkdir -p gpurun_out
python tools/profile_step.py 32 xl 384 > gpurun_out/profile_step_xl.log 2>&1; echo "profile rc $?"; head -45 gpurun_out/profile_step_xl.log
