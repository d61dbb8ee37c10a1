set -x
python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8 > gpurun_out/r2_tests_final5.log; tail -3 gpurun_out/r2_tests_final5.log
timeout 1500 python bench.py > gpurun_out/bench_r02_final5.json 2> gpurun_out/bench_r02_final5.err; echo bench rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
