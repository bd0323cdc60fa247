#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --mode infer --steps 10 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_infer_b256.log 2>&1; echo "infer rc $?"; tail -1 gpurun_out/bench_infer_b256.log | cut -c1-500
timeout 600 python bench.py --variant md --steps 3 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_md_b256.log 2>&1; echo "md rc $?"; tail -1 gpurun_out/bench_md_b256.log | cut -c1-300
timeout 600 python bench.py --variant xl --img 384 --steps 3 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench_xl_b32.log 2>&1; echo "xl rc $?"; tail -2 gpurun_out/bench_xl_b32.log | cut -c1-400
