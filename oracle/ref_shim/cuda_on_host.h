// cuda_on_host.h -- ORACLE infrastructure (tests only, never product code).
// The handful of CUDA names the reference's kernel headers (Non-CDP/GASAL2/src/gasal_kernels.h and
// src/kernels/*.h) use, defined so that those headers compile UNMODIFIED with g++ and every
// __global__ template becomes an ordinary function run once per "thread" (SURVEY.md Appendix A.2).
// Nothing from the reference is copied here: the reference sources are #included from where they
// lie under /root/reference at build time (oracle/Makefile), and only the resulting .so is kept
// (oracle/_ref/, git-ignored).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <limits.h>
#include <string.h>
#include <algorithm>
#include <math.h>

#define __global__
#define __device__
#define __host__
#define __constant__ static
#define __forceinline__ inline

struct shim_dim3 { unsigned x, y, z; };
static thread_local shim_dim3 blockIdx, blockDim, threadIdx, gridDim;

struct short2 { short x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline short2 make_short2(int a, int b) { short2 r; r.x = (short)a; r.y = (short)b; return r; }
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 r = {a, b, c, d}; return r; }
using std::max;
using std::min;

// the reference's own result struct and enums (src/gasal.h:37-95) -- gasal.h itself cannot be
// included because it pulls in cuda_runtime.h through an absolute path.
struct gasal_res {
  int32_t *aln_score, *query_batch_end, *target_batch_end, *query_batch_start, *target_batch_start;
  uint8_t *cigar;
  uint32_t *n_cigar_ops;
};
typedef struct gasal_res gasal_res_t;
enum comp_start { WITHOUT_START, WITH_START, WITH_TB };
enum Bool { FALSE, TRUE };
enum data_source { NONE, QUERY, TARGET, BOTH };
enum algo_type { UNKNOWN, GLOBAL, SEMI_GLOBAL, LOCAL, MICROLOCAL, BANDED, KSW };
