for v in main risky99 noprune_rev; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"; python tools/bench_real.py 2>&1 | tail -5
done > gpurun_out/r2_real11.log 2>&1
cat gpurun_out/r2_real11.log
unset GASAL_B200_LIB
python -m pytest tests -m gpu -q --maxfail=20 -k "parity_holes or parity_gpu or ref_gpu" 2>&1 | tail -4
