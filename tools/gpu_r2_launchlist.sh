set -x
python bench.py --steps 2 --warmup 1 --extras 0 --refgpu-pairs 0 --cpu-pairs 1000 > gpurun_out/r2_bench_short_plain.json 2> gpurun_out/r2_bench_short_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/ncu_launches_r02.csv python bench.py --steps 2 --warmup 1 --extras 0 --refgpu-pairs 0 --cpu-pairs 1000 > gpurun_out/r2_bench_short_ncu.json 2> gpurun_out/r2_bench_short_ncu.err
echo rc=$?; wc -l gpurun_out/ncu_launches_r02.csv
