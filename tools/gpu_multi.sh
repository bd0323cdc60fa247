#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/bench_dp$N.log 2>&1
echo "dp$N rc $?"; tail -3 gpurun_out/bench_dp$N.log | cut -c1-600
