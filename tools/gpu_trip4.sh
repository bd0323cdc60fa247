#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_attn_tc.py tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py tests/test_gpu_model.py
timeout 600 python bench.py --steps 5 --warmup 2 --batch 256 --no-cpu-baseline > gpurun_out/bench_graph_b256.log 2>&1; echo "graph b256 rc $?"; tail -1 gpurun_out/bench_graph_b256.log | cut -c1-300
LNX_GEMM_V1=1 timeout 600 python bench.py --steps 5 --warmup 2 --batch 256 --no-cpu-baseline > gpurun_out/bench_graph_b256_v1.log 2>&1; echo "v1 rc $?"; tail -1 gpurun_out/bench_graph_b256_v1.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 1 --batch 256 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 3500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --batch 256 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"
