#!/bin/bash
# A/B of the committed tree (_ab_old/, built by hand) against the working tree, interleaved, same box.
# Usage (on the GPU box): bash tools/ab.sh [batch] [steps]
B=${1:-4}; S=${2:-50}
for i in 1 2 3; do
  for t in _ab_old .; do
    (cd $t && python bench.py --steps $S --warmup 5 --batch $B --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$t', 'batch', $B, 'ms/step', round(d['ms_per_step'],3), 'conv TF/s', round(d['roofline']['achieved']))")
  done
done
