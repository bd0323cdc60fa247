#!/bin/bash
# Same-box A/B: _ab_base/ (a built copy of an earlier commit) against the working tree, alternating, train and inference.
# usage (on the GPU box): tools/ab_bench.sh [rounds]
R=${1:-2}
one() {  # dir label args...
  (cd "$1" && shift && lbl=$1 && shift && python bench.py "$@" --no-sub --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lbl', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))")
}
for i in $(seq $R); do
  one _ab_base "base train" --steps 20 --warmup 5
  one . "new  train" --steps 20 --warmup 5
  one _ab_base "base infer" --mode infer --steps 30 --warmup 5
  one . "new  infer" --mode infer --steps 30 --warmup 5
done
