set -x
python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -30 > gpurun_out/r2_tests_final.log; tail -6 gpurun_out/r2_tests_final.log
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python tools/bench_c4.py 262144 > gpurun_out/r2_c4_final.log 2>&1; cat gpurun_out/r2_c4_final.log
