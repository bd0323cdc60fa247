#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py tests/test_gpu_model.py 2>&1 | grep -E "exit|passed|failed|^E  " | head -20
python tools/prof_small_gemm.py 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_graph_b256.log 2>&1; echo "graph b256 rc $?"; tail -1 gpurun_out/bench_graph_b256.log | cut -c1-300
