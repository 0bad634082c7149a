#!/bin/bash
# Same-box comparison of several environment settings of the step: each argument is one setting, e.g.
#   bash tools/ab_envs.sh "UB2_WGRAD_SIDE=0" "UB2_WGRAD_SIDE=1 UB2_WGRAD_SMEM_KB=192"
# Interleaved, three rounds (box-to-box spread on this pool is +-1.5 %).
B=${BATCH:-4}; S=${STEPS:-40}
for i in 1 2 3; do
  for v in "$@"; do
    env $v python "$(dirname "$0")/../bench.py" --steps $S --warmup 5 --batch $B --no-cpu-baseline --no-extras 2>>gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'batch', $B, 'ms/step', round(d['ms_per_step'],3), 'dp_check', d.get('dp_check'))"
  done
done
