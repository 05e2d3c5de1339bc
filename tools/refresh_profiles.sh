# The command sequence behind profiles/r01d_*: run on the GPU box through gpurun, after `bench.py` has exited 0 without ncu.
set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_d.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair_tmem -s 4 -c 1 -f -o gpurun_out/prof_d_pair_tmem python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_d_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hourglass1 -s 4 -c 1 -f -o gpurun_out/prof_d_hourglass1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_d_hourglass.log 2>&1
ls -la gpurun_out/*_d*
