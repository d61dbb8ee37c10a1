set -x
python -m pytest tests -m gpu -q --maxfail=10 -k "parity_gpu or parity_holes or fullsize or host_packing" 2>&1 | tail -15 > gpurun_out/r2_tests2.log; tail -6 gpurun_out/r2_tests2.log
export QUICK_POLICIES=0
for v in main u4 u1 base; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"; python tools/quick_bench.py 262144 2>&1 | grep -v "^$"
done > gpurun_out/r2_quick2.log 2>&1
cat gpurun_out/r2_quick2.log
