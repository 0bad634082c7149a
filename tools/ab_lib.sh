#!/bin/bash
# Same-box A/B of two builds of libunetb200.so (run on the GPU box, through gpurun):
#   python unet-segment-pytorch_b200/build.py --variant old SOME_DEFINE   ->  lib/libunetb200.old.so
#   bash tools/ab_lib.sh old [batch] [steps]
# Interleaves bench.py runs of the variant (UB2_LIB) and the default build; box-to-box spread on this pool is
# +-1.5 %, so only interleaved runs on one box compare kernels.
V=${1:?variant name}; B=${2:-4}; S=${3:-40}
L=$(dirname "$0")/../unet-segment-pytorch_b200/lib/libunetb200.$V.so
for i in 1 2 3; do
  for lib in "$L" ""; do
    UB2_LIB=$lib python "$(dirname "$0")/../bench.py" --steps $S --warmup 5 --batch $B --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('${lib:-default}'.split('/')[-1], 'batch', $B, 'ms/step', round(d['ms_per_step'],3), 'conv TF/s', round(d['roofline']['achieved']))"
  done
done
