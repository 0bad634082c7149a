set -x
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -o gpurun_out/r2_prof_conv_b4 -f python tools/prof_conv.py 4 > gpurun_out/r2_prof_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 900 ncu --set full --clock-control none -o gpurun_out/r2_prof_mem_b4 -f python tools/prof_mem.py 4 > gpurun_out/r2_prof_mem.log 2>&1; echo "ncu mem rc=$?"
tail -4 gpurun_out/r2_gputest.log
