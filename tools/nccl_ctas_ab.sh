#!/bin/bash
# Same-box A/B of NCCL_MAX_CTAS on the 4-GPU train step (does capping the CTAs NCCL may take from the persistent compute kernels help?).
# Measured (round 2): default 18.38 ms, 8 -> 18.47 ms, 4 -> 18.71 ms per step: no; the default stays.
for v in default 8 4; do
  if [ "$v" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-sub 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NCCL_MAX_CTAS=$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
