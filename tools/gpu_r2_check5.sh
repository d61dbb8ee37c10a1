set -x
python -m pytest tests -m gpu -q --maxfail=20 -k "host_packing or parity_holes or parity_gpu or wave_gpu or banded or mixed or fullsize" 2>&1 | tail -30 > gpurun_out/r2_tests5.log; tail -12 gpurun_out/r2_tests5.log
export QUICK_POLICIES=0
python tools/quick_bench.py 262144 > gpurun_out/r2_quick5.log 2>&1; cat gpurun_out/r2_quick5.log
python tools/bench_c4.py 262144 2>&1 | head -3
