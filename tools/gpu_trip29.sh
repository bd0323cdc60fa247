#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph_b256.log 2>&1; echo "train rc $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_graph_b256.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['clocks'])
PY
timeout 300 python bench.py --mode infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_b256.log 2>&1; echo "infer rc $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_infer_b256.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
PY
timeout 300 python bench.py --arch v0 --mode infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v0_infer_b256.log 2>&1; echo "v0 rc $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v0_infer_b256.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'])
PY
