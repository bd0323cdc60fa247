#!/bin/bash
mkdir -p gpurun_out
python tools/profile_step.py 256 > gpurun_out/profile_step_b256.log 2>&1; echo "profile rc $?"; head -75 gpurun_out/profile_step_b256.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_graph_b256.log 2>&1; echo "graph b256 rc $?"; tail -1 gpurun_out/bench_graph_b256.log | cut -c1-400
