#!/bin/bash
# Round-end validation: every GPU test file, smoke(), the default bench (with CPU baseline), the reference arm, inference and
# V0 benches, the metrics measurement.
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_attn_tc.py tests/test_gpu_droppath.py tests/test_gpu_model.py tests/test_gpu_v0.py tests/test_gpu_metrics.py tests/test_gpu_aug.py 2>&1 | grep -E "exit|passed|failed|Error|error"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench rc $?"; tail -1 gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference rc $?"; tail -1 gpurun_out/bench_reference.log
timeout 300 python bench.py --mode infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_b256.log 2>&1; echo "infer rc $?"; tail -1 gpurun_out/bench_infer_b256.log | cut -c1-400
timeout 300 python bench.py --mode infer --arch v0 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0_infer_b256.log 2>&1; echo "v0 infer rc $?"; tail -1 gpurun_out/bench_v0_infer_b256.log | cut -c1-400
timeout 300 python tools/prof_metrics.py > gpurun_out/prof_metrics.log 2>&1; echo "metrics rc $?"; cat gpurun_out/prof_metrics.log
timeout 300 python bench.py --variant md --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_md_b256.log 2>&1; echo "md rc $?"; tail -1 gpurun_out/bench_md_b256.log | cut -c1-200
timeout 400 python bench.py --variant xl --img 384 --batch 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_xl_b32.log 2>&1; echo "xl rc $?"; tail -1 gpurun_out/bench_xl_b32.log | cut -c1-200
timeout 300 python bench.py --mode infer --arch v0 --batch 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v0_infer_b1.log 2>&1; echo "v0 b1 rc $?"; tail -1 gpurun_out/bench_v0_infer_b1.log | cut -c1-200
timeout 300 python bench.py --mode infer --arch v0 --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0_infer_b1024.log 2>&1; echo "v0 b1024 rc $?"; tail -1 gpurun_out/bench_v0_infer_b1024.log | cut -c1-200
timeout 300 python tools/prof_aug.py > gpurun_out/prof_aug.log 2>&1; echo "aug rc $?"; cat gpurun_out/prof_aug.log
