#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"dwconv7_fwd_x2|dwconv7_wgrad_x2|ln_bwd_bf16" -s 6 -c 4 -o gpurun_out/prof_ops2 -f python tools/prof_ops.py 256 > gpurun_out/ncu_ops2.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_ops2.log
