#!/bin/bash
# every GPU test file + step profile + graph bench
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_attn_tc.py tests/test_gpu_droppath.py tests/test_gpu_model.py tests/test_gpu_v0.py tests/test_gpu_metrics.py tests/test_gpu_aug.py 2>&1 | grep -E "exit|passed|failed|Error|error" 
python tools/profile_step.py 256 > gpurun_out/profile_step_b256.log 2>&1; echo "profile rc $?"; head -30 gpurun_out/profile_step_b256.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_graph_b256.log 2>&1; echo "graph b256 rc $?"; tail -1 gpurun_out/bench_graph_b256.log | cut -c1-300
