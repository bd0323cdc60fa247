#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke.log
bash tools/run_gpu_checks.sh tests/test_gpu_v0.py 2>&1 | grep -E "exit|passed|failed|^E  " | head
python tools/profile_v0.py 256 > gpurun_out/profile_v0_b256.log 2>&1; head -50 gpurun_out/profile_v0_b256.log
