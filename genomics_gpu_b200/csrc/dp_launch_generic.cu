// dp_launch_generic.cu -- translation unit of the general int32 kernels (dp_generic.cuh), the pack / reverse-complement kernels,
// the traceback walker and the CIGAR text kernel (pack_tb.cuh).  Dispatch replaces the reference's macro tower
// (Non-CDP/GASAL2/src/gasal_align.h:7-107): head / tail and the scoring are run-time kernel arguments.
#include <algorithm>
#include "dp_launch.h"
#include "dp_generic.cuh"
#include "pack_tb.cuh"

namespace gb200 {

int generic_launch(const DpArgs& a, int algo, int start, bool second, cudaStream_t str) {
  const uint32_t n = (uint32_t)a.n, BLOCKDIM = 128, N_BLOCKS = (n + BLOCKDIM - 1) / BLOCKDIM;
  int launches = 1;
  switch (algo) {
    case ALGO_LOCAL:
      if (start == START_TB) { if (second) k_local_generic<START_TB, true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); else k_local_generic<START_TB, false><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); }
      else if (start == START_POS) { if (second) k_local_generic<START_POS, true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); else k_local_generic<START_POS, false><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); }
      else { if (second) k_local_generic<START_NONE, true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); else k_local_generic<START_NONE, false><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); }
      if (start == START_TB) { k_get_tb<true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); launches++; }   // gasal_align.h:45-53
      break;
    case ALGO_SEMI_GLOBAL:
      if (start == START_POS) { if (second) k_semiglobal_generic<START_POS, true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); else k_semiglobal_generic<START_POS, false><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); }
      else { if (second) k_semiglobal_generic<START_NONE, true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); else k_semiglobal_generic<START_NONE, false><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); }
      break;
    case ALGO_GLOBAL:
      if (start == START_TB) k_global_generic<START_TB><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); else k_global_generic<START_NONE><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a);
      if (start == START_TB) { k_get_tb<false><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); launches++; }   // gasal_align.h:60-68
      break;
    case ALGO_BANDED:
      if (start == START_TB) {
        k_banded_generic<true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a);
        k_get_tb<true><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a); launches++;
      } else k_banded_generic<false><<<N_BLOCKS, BLOCKDIM, 0, str>>>(a);
      break;
    case ALGO_KSW:                                                                          // gasal_align.h:73-78
      k_ksw_generic<<<N_BLOCKS, BLOCKDIM, 0, str>>>(a);
      break;
    case ALGO_KSW_SMEM: {                                                                   // row in shared memory (engine.cu proves the 16-bit range)
      const size_t smem = (size_t)(a.col_rows + 2) * kKswThreads * sizeof(uint32_t);
      static bool attr_done[64] = {};
      int dev = 0; cudaGetDevice(&dev);
      if (!attr_done[dev & 63]) { cudaFuncSetAttribute(k_ksw_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_done[dev & 63] = true; }
      k_ksw_smem<<<(n + kKswThreads - 1) / kKswThreads, kKswThreads, smem, str>>>(a);
      break;
    }
    default:
      return -1;
  }
  return launches;
}

int launch_pack(const uint8_t* uq, uint32_t* pq, uint32_t q_bytes, const uint8_t* ut, uint32_t* pt, uint32_t t_bytes, int n_sm, cudaStream_t str) {
  int launches = 0;
  const uint32_t n16q = q_bytes / 16, n16t = t_bytes / 16;
  const uint32_t total = n16q + n16t;
  if (total) {
    const uint32_t blocks = std::min<uint32_t>((total + 255) / 256, (uint32_t)n_sm * 8);
    k_pack<<<blocks, 256, 0, str>>>((const uint4*)uq, (uint2*)pq, n16q, (const uint4*)ut, (uint2*)pt, n16t);
    launches++;
  }
  if (q_bytes & 8) { k_pack_tail<<<1, 1, 0, str>>>((const uint2*)(uq + 16 * (size_t)n16q), pq + 2 * (size_t)n16q); launches++; }
  if (t_bytes & 8) { k_pack_tail<<<1, 1, 0, str>>>((const uint2*)(ut + 16 * (size_t)n16t), pt + 2 * (size_t)n16t); launches++; }
  return launches;
}

void launch_revcomp(uint32_t* pq, uint32_t* pt, const uint32_t* q_lens, const uint32_t* t_lens, const uint32_t* q_offs, const uint32_t* t_offs,
                    const uint8_t* q_op, const uint8_t* t_op, uint32_t n, int mode, cudaStream_t str) {
  const uint32_t warps = 2 * n, blocks = (warps * 32 + 255) / 256;
  k_revcomp<<<blocks, 256, 0, str>>>(pq, pt, q_lens, t_lens, q_offs, t_offs, q_op, t_op, n, mode);
}

void launch_get_tb(const DpArgs& a, bool is_local, cudaStream_t str) {
  const uint32_t nb = ((uint32_t)a.n + 127) / 128;
  if (is_local) k_get_tb<true><<<nb, 128, 0, str>>>(a); else k_get_tb<false><<<nb, 128, 0, str>>>(a);
}

struct ScatterArgs { const int32_t* src[5]; int32_t* dst[5]; };
__global__ void __launch_bounds__(256) k_scatter_results(const uint32_t* __restrict__ order, uint32_t n, const ScatterArgs sa) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = order[i];
#pragma unroll
  for (int k = 0; k < 5; k++) if (sa.src[k] && sa.dst[k]) sa.dst[k][p] = sa.src[k][i];
}

void launch_scatter_results(const uint32_t* order, uint32_t n, int32_t* const src[5], int32_t* const dst[5], cudaStream_t str) {
  ScatterArgs sa;
  for (int k = 0; k < 5; k++) { sa.src[k] = src[k]; sa.dst[k] = dst[k]; }
  k_scatter_results<<<(n + 255) / 256, 256, 0, str>>>(order, n, sa);
}

void launch_cigar_text(const uint8_t* cigar, const uint32_t* n_cigar_ops, const uint32_t* q_offs, const uint32_t* q_lens, const int32_t* q_start,
                       const int32_t* q_end, int n, int style, int clip, char* text, uint32_t stride, uint32_t* lens, cudaStream_t str) {
  k_cigar_text<<<(n + 127) / 128, 128, 0, str>>>(cigar, n_cigar_ops, q_offs, q_lens, q_start, q_end, n, style, clip, text, stride, lens);
}

}  // namespace gb200
