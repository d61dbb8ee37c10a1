set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q --maxfail=15 -x -k "host_packing or parity_holes or nvbio or parity_gpu" 2>&1 | tail -40 > gpurun_out/r2_tests_new.log; tail -15 gpurun_out/r2_tests_new.log
python -m pytest tests -m gpu -q --maxfail=15 2>&1 | tail -25 > gpurun_out/r2_tests_all.log; tail -8 gpurun_out/r2_tests_all.log
timeout 900 python bench.py --steps 10 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench1.err
python tools/quick_bench.py 262144 > gpurun_out/r2_quick1.log 2>&1; cat gpurun_out/r2_quick1.log
