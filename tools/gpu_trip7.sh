#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_model.py
python tools/prof_gemm.py 256 > gpurun_out/prof_gemm.log 2>&1 && cat gpurun_out/prof_gemm.log && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 6 -c 3 -o gpurun_out/prof_gemm2 python tools/prof_gemm.py 256 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc $?"
timeout 600 python bench.py --steps 5 --warmup 2 --batch 256 --no-cpu-baseline > gpurun_out/bench_graph_b256.log 2>&1; echo "graph b256 rc $?"; tail -1 gpurun_out/bench_graph_b256.log | cut -c1-300
