#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dwconv7 -s 4 -c 4 -o gpurun_out/prof_dwconv -f python tools/prof_ops.py 256 > gpurun_out/ncu_dwconv.log 2>&1
echo "ncu rc $?"; tail -3 gpurun_out/ncu_dwconv.log
