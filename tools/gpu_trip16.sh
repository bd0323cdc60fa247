#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc|gemm_tc_kernel" -s 4 -c 2 -o gpurun_out/prof_wgrad -f python tools/prof_wgrad.py > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu rc $?"; tail -3 gpurun_out/ncu_wgrad.log
