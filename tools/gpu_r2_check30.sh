for v in main mb2 mb3; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"; QUICK_POLICIES=0 python tools/quick_bench.py 2>&1 | cut -c1-20,60-150
done
