# The command sequence behind profiles/r02_*: run on the GPU box through gpurun (one call), every ncu pass only after the
# same command exited 0 without ncu.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gemm-peak --no-global512 --no-fullpage --no-stages > gpurun_out/r02_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gemm-peak --no-global512 --no-fullpage --no-stages > gpurun_out/r02_ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-gemm-peak --no-fullpage --no-stages > gpurun_out/r02_plain_bench2.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair_rows -s 4 -c 1 -f -o gpurun_out/r02_pair_rows python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-gemm-peak --no-fullpage --no-stages > gpurun_out/r02_ncu_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hourglass1 -s 4 -c 1 -f -o gpurun_out/r02_hourglass1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-gemm-peak --no-fullpage --no-stages > gpurun_out/r02_ncu_hourglass.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hourglass4_fwd -s 4 -c 1 -f -o gpurun_out/r02_hourglass4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-gemm-peak --no-fullpage --no-stages > gpurun_out/r02_ncu_hourglass4.log 2>&1
timeout 300 python tools/microbench.py --math tf32 > gpurun_out/r02_microbench_tf32.txt 2>&1; tail -3 gpurun_out/r02_microbench_tf32.txt
ls -la gpurun_out/r02_*
