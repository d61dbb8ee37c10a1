set -x
python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8 > gpurun_out/r2_tests_final3.log; tail -4 gpurun_out/r2_tests_final3.log
timeout 1500 python bench.py > gpurun_out/bench_r02_final3.json 2> gpurun_out/bench_r02_final3.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02_final3.json 2> gpurun_out/bench_ref_r02_final3.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 --extras 0 --refgpu-pairs 0 --cpu-pairs 1000 > gpurun_out/r2_bench_short_plain.json 2> gpurun_out/r2_bench_short_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/ncu_launches_r02c.csv python bench.py --steps 2 --warmup 1 --extras 0 --refgpu-pairs 0 --cpu-pairs 1000 > gpurun_out/r2_bench_short_ncu.json 2> gpurun_out/r2_bench_short_ncu.err
echo ncu rc=$?; wc -l gpurun_out/ncu_launches_r02c.csv
