python -m pytest tests -m gpu -q --maxfail=20 -k "parity_holes or parity_gpu or fullsize or mixed or ref_gpu" 2>&1 | tail -4
export QUICK_POLICIES=0
for v in main nolow norev; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"; python tools/quick_bench.py 262144 2>&1 | grep "local"
done
