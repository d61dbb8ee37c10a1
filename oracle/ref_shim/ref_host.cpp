// ref_host.cpp -- ORACLE infrastructure (tests only, never product code).
// Runs the reference's own GASAL2 kernels (Non-CDP/GASAL2/src/kernels/*.h, included unmodified
// through cuda_on_host.h) on the CPU, one OpenMP iteration per CUDA thread, behind a small C ABI.
// Built by oracle/Makefile into oracle/_ref/libgasal_ref_host_mq<MAX_QUERY_LEN>.so.
// Not valid for gasal_reversecomplement_kernel (undefined shifts) -- not exported.
#include "cuda_on_host.h"
#ifndef MAX_QUERY_LEN
#error "build with -DMAX_QUERY_LEN=<n>"
#endif
#ifndef N_CODE
#define N_CODE 0x4E
#endif
#include "gasal_kernels.h"   // from /root/reference/Non-CDP/GASAL2/src (via -I)
#include <vector>
#include <omp.h>

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int ref_max_query_len(void) { return MAX_QUERY_LEN; }

REF_API void ref_set_scores(int match, int mismatch, int gap_open, int gap_extend) {
  _cudaMatchScore = match; _cudaMismatchScore = mismatch; _cudaGapO = gap_open; _cudaGapExtend = gap_extend;
  _cudaGapOE = gap_open + gap_extend;   // gasal_align.cu:329-339
}

// gasal_pack_kernel with the launch shape of gasal_align.cu:180-194
REF_API void ref_pack(const uint8_t* uq, const uint8_t* ut, uint32_t* pq, uint32_t* pt, uint32_t q_bytes, uint32_t t_bytes, uint32_t n_alns) {
  const uint32_t BLOCKDIM = 128, N_BLOCKS = (n_alns + BLOCKDIM - 1) / BLOCKDIM;
  int qtpt = (int)ceil((double)q_bytes / (8 * BLOCKDIM * N_BLOCKS));
  int ttpt = (int)ceil((double)t_bytes / (8 * BLOCKDIM * N_BLOCKS));
  for (uint32_t b = 0; b < N_BLOCKS; b++)
    for (uint32_t t = 0; t < BLOCKDIM; t++) {
      blockIdx.x = b; blockDim.x = BLOCKDIM; threadIdx.x = t; gridDim.x = N_BLOCKS;
      gasal_pack_kernel((uint32_t*)uq, (uint32_t*)ut, pq, pt, qtpt, ttpt, q_bytes / 4, t_bytes / 4);
    }
}

template <typename S, typename B, typename H, typename T>
static void run_sg(uint32_t* pq, uint32_t* pt, uint32_t* ql, uint32_t* tl, uint32_t* qo, uint32_t* to, gasal_res_t* r, gasal_res_t* r2, int n) {
  gasal_semi_global_kernel<Int2Type<SEMI_GLOBAL>, S, B, H, T>(pq, pt, ql, tl, qo, to, r, r2, NULL, n);
}

template <typename S, typename B, typename H>
static void sg_tail(int tail, uint32_t* pq, uint32_t* pt, uint32_t* ql, uint32_t* tl, uint32_t* qo, uint32_t* to, gasal_res_t* r, gasal_res_t* r2, int n) {
  switch (tail) {
    case NONE: run_sg<S, B, H, Int2Type<NONE> >(pq, pt, ql, tl, qo, to, r, r2, n); break;
    case QUERY: run_sg<S, B, H, Int2Type<QUERY> >(pq, pt, ql, tl, qo, to, r, r2, n); break;
    case TARGET: run_sg<S, B, H, Int2Type<TARGET> >(pq, pt, ql, tl, qo, to, r, r2, n); break;
    default: run_sg<S, B, H, Int2Type<BOTH> >(pq, pt, ql, tl, qo, to, r, r2, n); break;
  }
}
template <typename S, typename B>
static void sg_head(int head, int tail, uint32_t* pq, uint32_t* pt, uint32_t* ql, uint32_t* tl, uint32_t* qo, uint32_t* to, gasal_res_t* r, gasal_res_t* r2, int n) {
  switch (head) {
    case NONE: sg_tail<S, B, Int2Type<NONE> >(tail, pq, pt, ql, tl, qo, to, r, r2, n); break;
    case QUERY: sg_tail<S, B, Int2Type<QUERY> >(tail, pq, pt, ql, tl, qo, to, r, r2, n); break;
    case TARGET: sg_tail<S, B, Int2Type<TARGET> >(tail, pq, pt, ql, tl, qo, to, r, r2, n); break;
    default: sg_tail<S, B, Int2Type<BOTH> >(tail, pq, pt, ql, tl, qo, to, r, r2, n); break;
  }
}

// One launch of the reference's alignment kernel (+ gasal_get_tb for LOCAL/GLOBAL WITH_TB), the CPU
// analogue of gasal_kernel_launcher (gasal_align.cu:10-25 / gasal_align.h:7-107). All arrays are
// host arrays with the layout gasal_aln_async gives the device.  `cigar` must be q_bytes long.
REF_API int ref_align(int algo, int start, int second_best, int head, int tail, int k_band,
                      uint32_t* pq, uint32_t* pt, uint32_t* ql, uint32_t* tl, uint32_t* qo, uint32_t* to, uint32_t* seed, int n,
                      int32_t* score, int32_t* q_end, int32_t* t_end, int32_t* q_start, int32_t* t_start,
                      int32_t* score2, int32_t* q_end2, int32_t* t_end2,
                      uint8_t* cigar, uint32_t* n_cigar_ops, int n_threads) {
  gasal_res_t res = {score, q_end, t_end, q_start, t_start, cigar, n_cigar_ops};
  gasal_res_t res2 = {score2, q_end2, t_end2, NULL, NULL, NULL, NULL};
  std::vector<uint4> tb;
  std::vector<uint32_t> ql_copy;
  if (start == WITH_TB && (algo == LOCAL || algo == GLOBAL)) {
    uint64_t mq = 0, mt = 0;
    for (int i = 0; i < n; i++) { uint64_t a = (ql[i] + 7) / 8 * 8, b = (tl[i] + 7) / 8 * 8; mq = std::max(mq, a); mt = std::max(mt, b); }
    tb.assign(((mq * (mt + 8) + 31) / 32 + 2) * (uint64_t)n, make_uint4(0, 0, 0, 0));
    ql_copy.assign(ql, ql + n);
  }
  if (n_threads > 0) omp_set_num_threads(n_threads);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (int tid = 0; tid < n; tid++) {
    blockIdx.x = tid; blockDim.x = 1; threadIdx.x = 0; gridDim.x = n;
#define LOC(S, B) gasal_local_kernel<Int2Type<LOCAL>, Int2Type<S>, Int2Type<B> >(pq, pt, ql, tl, qo, to, &res, &res2, tb.data(), n)
#define SG(S, B) sg_head<Int2Type<S>, Int2Type<B> >(head, tail, pq, pt, ql, tl, qo, to, &res, &res2, n)
    switch (algo) {
      case LOCAL:
        if (start == WITH_TB) { if (second_best) LOC(WITH_TB, TRUE); else LOC(WITH_TB, FALSE); }
        else if (start == WITH_START) { if (second_best) LOC(WITH_START, TRUE); else LOC(WITH_START, FALSE); }
        else { if (second_best) LOC(WITHOUT_START, TRUE); else LOC(WITHOUT_START, FALSE); }
        break;
      case SEMI_GLOBAL:
        if (start == WITH_START) { if (second_best) SG(WITH_START, TRUE); else SG(WITH_START, FALSE); }
        else { if (second_best) SG(WITHOUT_START, TRUE); else SG(WITHOUT_START, FALSE); }
        break;
      case GLOBAL:
        if (start == WITH_TB) gasal_global_kernel<Int2Type<WITH_TB> >(pq, pt, ql, tl, qo, to, &res, tb.data(), n);
        else gasal_global_kernel<Int2Type<WITHOUT_START> >(pq, pt, ql, tl, qo, to, &res, tb.data(), n);
        break;
      case BANDED:
        gasal_banded_tiled_kernel(pq, pt, ql, tl, qo, to, &res, n, k_band >> 3);   // gasal_align.h:80
        break;
      case KSW:
        if (second_best) gasal_ksw_kernel<Int2Type<TRUE> >(pq, pt, ql, tl, qo, to, seed, &res, &res2, n);
        else gasal_ksw_kernel<Int2Type<FALSE> >(pq, pt, ql, tl, qo, to, seed, &res, &res2, n);
        break;
      default: bad = 1; break;
    }
  }
  if (bad) return -1;
  if (start == WITH_TB && (algo == LOCAL || algo == GLOBAL)) {
    // second launch, gasal_align.h:45-53,60-68: cigar -> unpacked query buffer, n_ops -> query lens
#pragma omp parallel for schedule(dynamic, 16)
    for (int tid = 0; tid < n; tid++) {
      blockIdx.x = tid; blockDim.x = 1; threadIdx.x = 0; gridDim.x = n;
      if (algo == LOCAL) gasal_get_tb<Int2Type<LOCAL> >(cigar, ql_copy.data(), tl, qo, tb.data(), &res, n);
      else gasal_get_tb<Int2Type<GLOBAL> >(cigar, ql_copy.data(), tl, qo, tb.data(), &res, n);
    }
    for (int i = 0; i < n; i++) n_cigar_ops[i] = ql_copy[i];
  }
  return 0;
}
