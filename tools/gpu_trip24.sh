#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_checks.sh tests/test_gpu_v0.py 2>&1 | tail -40
