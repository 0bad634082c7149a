# Final measurements of a round on one B200 (run through gpurun; every output is small: the ncu reports stay
# on the box, only their summaries come back).  bash tools/final_gpu.sh [tests|bench|launches|ncu]...
set -u
T0=$(date +%s); lap() { echo "== $1: $(( $(date +%s) - T0 )) s"; }
for what in "$@"; do case $what in
tests)
  timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest.log
  tail -4 gpurun_out/r2_gputest.log; lap tests;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; lap bench;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b4.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
  python tools/launch_summary.py gpurun_out/r2_launches_b4.csv > gpurun_out/r2_launches_b4.md 2>&1
  # keep only one step of the launch list (the csv of the whole run is tens of MB)
  python - <<'PY'
import re
lines = open("gpurun_out/r2_launches_b4.csv").read().splitlines()
hdr = [i for i, l in enumerate(lines) if l.startswith('"ID"')]
body = lines[hdr[0] + 1:] if hdr else lines
marks = [i for i, l in enumerate(body) if "pack_weights_multi" in l]
keep = body[marks[-2]:marks[-1]] if len(marks) >= 2 else body[-300:]
open("gpurun_out/r2_launches_b4.csv", "w").write("\n".join(([lines[hdr[0]]] if hdr else []) + keep) + "\n")
PY
  lap launches;;
ncu)
  timeout 900 ncu --set full --clock-control none -k regex:conv_ --launch-skip 18 --launch-count 18 -o /tmp/r2_prof_conv_b4 -f \
      python tools/prof_conv.py 4 > gpurun_out/r2_prof_conv.log 2>&1; echo "ncu conv rc=$?"
  python tools/ncu_summary.py /tmp/r2_prof_conv_b4.ncu-rep > gpurun_out/r2_ncu_conv_b4.md 2>&1
  python tools/conv_traffic.py /tmp/r2_prof_conv_b4.ncu-rep 4 > gpurun_out/r2_conv_traffic.json 2>> gpurun_out/r2_prof_conv.log
  lap ncu-conv
  timeout 900 ncu --set full --clock-control none -k 'regex:bn_|gate_|upsample|outc_|conv_in|seg_stats|shuffle' --launch-skip 20 -o /tmp/r2_prof_mem_b4 -f \
      python tools/prof_mem.py 4 > gpurun_out/r2_prof_mem.log 2>&1; echo "ncu mem rc=$?"
  python tools/ncu_summary.py /tmp/r2_prof_mem_b4.ncu-rep > gpurun_out/r2_ncu_mem_b4.md 2>&1
  lap ncu-mem;;
esac; done
du -sh gpurun_out | tail -1
