set -x
python -m pytest tests -m gpu -q --maxfail=20 -k "parity_holes or parity_gpu or fullsize or host_packing or mixed" 2>&1 | tail -30 > gpurun_out/r2_tests8.log; tail -8 gpurun_out/r2_tests8.log
export QUICK_POLICIES=0
python tools/quick_bench.py 262144 2>&1 | grep "local" > gpurun_out/r2_quick8.log 2>&1
cat gpurun_out/r2_quick8.log
