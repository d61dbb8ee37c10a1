set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_tests_final2.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err
tail -3 gpurun_out/r2_tests_final2.log
