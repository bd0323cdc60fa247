#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_v0.py 2>&1 | grep -E "exit|passed|failed|^E  " | head
python tools/profile_v0.py 256 2>&1 | grep -E "forward|attn_bias"
timeout 300 python bench.py --mode infer --arch v0 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0_infer_b256.log 2>&1; echo "v0 infer rc $?"; tail -1 gpurun_out/bench_v0_infer_b256.log | cut -c1-200
