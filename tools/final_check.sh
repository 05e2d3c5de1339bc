set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/microbench.py --math tf32 > gpurun_out/microbench_tf32_d.txt 2>&1; tail -3 gpurun_out/microbench_tf32_d.txt
timeout 400 python bench.py > gpurun_out/bench_r1_final_e.log 2> gpurun_out/bench_r1_final_e.err; tail -c 300 gpurun_out/bench_r1_final_e.log; tail -2 gpurun_out/bench_r1_final_e.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final_e.log 2>&1; tail -c 400 gpurun_out/bench_ref_final_e.log
