python -m pytest tests -m gpu -q --maxfail=20 -k "parity_holes or parity_gpu or fullsize or mixed or ref_gpu or host_packing or dropin" 2>&1 | tail -6
for v in main nolow; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"; python tools/bench_real.py 2>&1 | tail -4
done
