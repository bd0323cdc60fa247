#!/bin/bash
# Round-end validation: every GPU test file, smoke(), the default bench (with CPU baseline), the reference arm, inference and
# V0 benches, the metrics measurement.
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py tests/test_gpu_attn_tc.py tests/test_gpu_droppath.py tests/test_gpu_model.py tests/test_gpu_v0.py tests/test_gpu_metrics.py tests/test_gpu_aug.py tests/test_gpu_bench_shapes.py tests/test_gpu_loss_semantics.py tests/test_gpu_mlp_fused.py tests/test_gpu_qkv_rope.py tests/test_gpu_postprocess.py tests/test_gpu_dwconv_mma.py 2>&1 | grep -E "exit|passed|failed|Error|error"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench rc $?"; tail -1 gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference rc $?"; tail -1 gpurun_out/bench_reference.log
timeout 300 python bench.py --mode infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_b256.log 2>&1; echo "infer rc $?"; tail -1 gpurun_out/bench_infer_b256.log | cut -c1-400
timeout 300 python bench.py --mode infer --arch v0 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0_infer_b256.log 2>&1; echo "v0 infer rc $?"; tail -1 gpurun_out/bench_v0_infer_b256.log | cut -c1-400
timeout 300 python tools/prof_metrics.py > gpurun_out/prof_metrics.log 2>&1; echo "metrics rc $?"; cat gpurun_out/prof_metrics.log
timeout 300 python tools/prof_aug.py > gpurun_out/prof_aug.log 2>&1; echo "aug rc $?"; cat gpurun_out/prof_aug.log
