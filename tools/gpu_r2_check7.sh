set -x
python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -30 > gpurun_out/r2_tests7.log; tail -8 gpurun_out/r2_tests7.log
timeout 1200 python bench.py --steps 10 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench7.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench7_ref.json 2> gpurun_out/r2_bench7_ref.err; echo ref rc=$?
python tools/bench_c4.py 262144 > gpurun_out/r2_c4_7.log 2>&1; cat gpurun_out/r2_c4_7.log
