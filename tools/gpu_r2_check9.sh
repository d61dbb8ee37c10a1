set -x
python -m pytest tests -m gpu -q --maxfail=20 -k "parity_holes or parity_gpu or fullsize or host_packing or mixed or ref_gpu" 2>&1 | tail -30 > gpurun_out/r2_tests9.log; tail -8 gpurun_out/r2_tests9.log
export QUICK_POLICIES=0
python tools/quick_bench.py 262144 2>&1 > gpurun_out/r2_quick9.log 2>&1
cat gpurun_out/r2_quick9.log
