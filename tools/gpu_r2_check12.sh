for v in main tb2 tbb3; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"; python tools/bench_c4.py 262144 2>&1 | grep CIGAR
done > gpurun_out/r2_c4_12.log 2>&1
cat gpurun_out/r2_c4_12.log
