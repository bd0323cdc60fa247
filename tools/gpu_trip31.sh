#!/bin/bash
mkdir -p gpurun_out
python tools/prof_small_gemm.py 2>&1 | tail -8
ncu --set full --clock-control none -k regex:gemm_simt -c 6 -o /tmp/simt -f python tools/prof_small_gemm.py > /dev/null 2>&1
ncu -i /tmp/simt.ncu-rep --page details 2>/dev/null | grep -E "gemm_simt|Duration|Registers|Executed Ipc|Local|Stall|One or More|L1/TEX Hit|Warp Cycles Per Issued|Issued Warp|DRAM Throughput|Max Bandwidth|Theoretical Occ|Achieved Occ" | head -80
