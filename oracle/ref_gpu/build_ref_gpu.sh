#!/bin/sh
# Build the REFERENCE Non-CDP GASAL2 library for sm_100 plus our driver (ref_gpu_driver.cpp) into ONE binary.
#   usage: build_ref_gpu.sh <path to reference GASAL2> <MAX_QUERY_LEN> <output binary>
# The reference sources are compiled from a scratch copy under /tmp (its header hard-codes an absolute
# cuda_runtime.h path that must be rewritten, exactly what its configure.sh does); nothing but the
# final binary is written into the repo (oracle/_ref/, git-ignored).  This is our own recipe -- the
# reference's Makefile/configure.sh are not run.
set -e
REF="$1"; MQ="$2"; OUT="$3"
HERE="$(cd "$(dirname "$0")" && pwd)"
CUDA=${CUDA_HOME:-/usr/local/cuda}
TMP="$(mktemp -d /tmp/ref_gpu_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF/src" "$TMP/src"
chmod -R u+w "$TMP/src"
sed -i "s#^\#include \".*cuda_runtime.h\"#\#include \"$CUDA/include/cuda_runtime.h\"#" "$TMP/src/gasal.h"
DEFS="-DMAX_QUERY_LEN=$MQ -DN_CODE=0x4E"
echo "[ref_gpu] nvcc gasal_align.cu (MAX_QUERY_LEN=$MQ, sm_100) ..."
"$CUDA/bin/nvcc" -c -O3 -std=c++11 -Xcompiler -w $DEFS --gpu-architecture=compute_100 --gpu-code=sm_100 -lineinfo \
    --default-stream per-thread "$TMP/src/gasal_align.cu" -o "$TMP/gasal_align.o"
for f in args_parser host_batch ctors interfaces res; do
  g++ -c -O3 -std=c++11 -w $DEFS "$TMP/src/$f.cpp" -o "$TMP/$f.o"
done
g++ -O3 -std=c++11 -w $DEFS -I"$TMP/src" -I"$CUDA/include" -c "$HERE/ref_gpu_driver.cpp" -o "$TMP/driver.o"
g++ -O3 -o "$OUT" "$TMP/driver.o" "$TMP/gasal_align.o" "$TMP/args_parser.o" "$TMP/host_batch.o" "$TMP/ctors.o" "$TMP/interfaces.o" "$TMP/res.o" \
    -L"$CUDA/lib64" -lcudart -lpthread
echo "[ref_gpu] built $OUT"
