#!/bin/bash
# build_variant.sh NAME "-DFLAG=1 ..."  ->  build/variants/NAME/libgasal_b200.so  (A/B builds for GPU experiments; GASAL_B200_LIB selects one)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/build/variants/$1"
make -s -C "$ROOT/genomics_gpu_b200/csrc" -j8 OUT="$ROOT/build/variants/$1/libgasal_b200.so" OBJDIR="$ROOT/build/variants/$1/obj" EXTRA="$2" 2>&1 | grep -i "error" || true
ls -la "$ROOT/build/variants/$1/libgasal_b200.so"
