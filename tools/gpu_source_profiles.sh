#!/bin/bash
# Source-level (per SASS line) stall samples of three hot kernels, for next round's kernel work.
mkdir -p gpurun_out
for k in "gemm_tc2_kernel<3" "dwconv7_fwd_x2" "attn_bwd_tc2"; do
  tag=$(echo "$k" | tr -cd 'a-z0-9_')
  timeout 300 ncu --set full --import-source on --clock-control none -k "regex:$k" -c 1 -o /tmp/src_$tag -f python tools/prof_kernels.py 256 1 > gpurun_out/ncu_src_$tag.log 2>&1
  echo "$tag rc $?"
  ncu -i /tmp/src_$tag.ncu-rep --page source --csv > gpurun_out/r01_source_$tag.csv 2>/dev/null
  ls -la gpurun_out/r01_source_$tag.csv
done
