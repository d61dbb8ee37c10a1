for v in main w2 w4; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"
  python tools/bench_long.py c3 128 2>&1 | tail -1
  python tools/bench_long.py l3 128 2>&1 | tail -1
  python tools/bench_long.py c5 4000 2>&1 | tail -1
done > gpurun_out/r2_long14.log 2>&1
cat gpurun_out/r2_long14.log
