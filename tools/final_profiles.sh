# round-end evidence: run on the GPU box (gpurun -- bash tools/final_profiles.sh); outputs under gpurun_out/, summarised into profiles/ by
#   python tools/summarize_profiles.py r02 gpurun_out/r02_launches.csv gpurun_out/r02_quad_bench.ncu-rep
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"quad_kernel|walk_kernel|mix_kernel|generic_kernel|rte_" -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r02_ncu_launch.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
