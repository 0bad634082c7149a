#!/bin/bash
# Same-box A/B of one environment knob of the library: bash tools/ab_env.sh UB2_UPSTATS_LOWRES 0 1 [batch] [steps]
K=${1:?env name}; A=${2:?value a}; Bv=${3:?value b}; B=${4:-4}; S=${5:-40}
for i in 1 2 3; do
  for v in "$A" "$Bv"; do
    env $K=$v python "$(dirname "$0")/../bench.py" --steps $S --warmup 5 --batch $B --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$K=$v', 'batch', $B, 'ms/step', round(d['ms_per_step'],3))"
  done
done
