#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_model.py
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/smoke.log
for b in 64 256; do
  timeout 600 python bench.py --steps 5 --warmup 2 --batch $b --no-graph --no-cpu-baseline > gpurun_out/bench_eager_b$b.log 2>&1; echo "eager b$b rc $?"; tail -2 gpurun_out/bench_eager_b$b.log
  timeout 600 python bench.py --steps 5 --warmup 2 --batch $b --no-cpu-baseline > gpurun_out/bench_graph_b$b.log 2>&1; echo "graph b$b rc $?"; tail -2 gpurun_out/bench_graph_b$b.log
done
timeout 600 python bench.py --steps 2 --warmup 1 --batch 256 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --batch 256 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc $?"
