#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --mode infer --steps 10 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_infer_b256.log 2>&1; echo "infer v1 rc $?"; tail -1 gpurun_out/bench_infer_b256.log | cut -c1-260
for b in 1 16 256 1024; do
timeout 300 python bench.py --arch v0 --mode infer --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench_v0_infer_b$b.log 2>&1; echo "v0 b$b rc $?"; tail -1 gpurun_out/bench_v0_infer_b$b.log | cut -c1-220
done
