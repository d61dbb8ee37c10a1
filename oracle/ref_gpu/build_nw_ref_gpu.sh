#!/bin/sh
# Builds the REFERENCE's NW benchmark (Non-CDP/NW: main.cu + kernel0..3.cu) for sm_100 with our own one-line recipe
# (not its Makefile), straight from the sources in place; only the binary lands in oracle/_ref/.  Used as the same-GPU
# baseline of the full-matrix NW kernel (tools/bench_nw.py) -- test/bench infrastructure, never linked into the product.
set -e
REF_NW="$1"; OUT="$2"; NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -gencode arch=compute_100,code=sm_100 -o "$OUT" "$REF_NW/main.cu" "$REF_NW/kernel0.cu" "$REF_NW/kernel1.cu" "$REF_NW/kernel2.cu" "$REF_NW/kernel3.cu"
