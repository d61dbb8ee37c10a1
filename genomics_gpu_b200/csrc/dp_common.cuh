// dp_common.cuh -- shared device helpers for the B200 alignment kernels.
// Packed-base layout is GASAL2's: 8 bases per uint32, first base in bits 31:28
// (reference Non-CDP/GASAL2/src/kernels/pack_rc_seqs.h:13-53).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace gb200 {

// enums mirror the reference's (src/gasal.h:37-72) so host code can pass them straight through
enum { START_NONE = 0, START_POS = 1, START_TB = 2 };
constexpr int START_REV = 3;      // kernel-internal: the start pass alone, on the forward results already in the result arrays (dp_s16.cuh)
enum { SKIP_NONE = 0, SKIP_QUERY = 1, SKIP_TARGET = 2, SKIP_BOTH = 3 };
enum { ALGO_GLOBAL = 1, ALGO_SEMI_GLOBAL = 2, ALGO_LOCAL = 3, ALGO_BANDED = 5, ALGO_KSW = 6 };
constexpr int ALGO_KSW_SMEM = 106;   // launcher-internal: KSW with the (H, E) row in shared memory (generic_launch; chosen by engine.cu)

constexpr int kMinusInf = -32768;   // MINUS_INF = SHRT_MIN, src/gasal_kernels.h:35

// Everything a DP kernel needs, passed by value as the kernel parameter (no process-global
// __constant__ state: the reference's _cudaGapO etc., src/gasal_kernels.h:29-33, become fields).
struct DpArgs {
  const uint32_t* __restrict__ pq;       // packed query batch
  const uint32_t* __restrict__ pt;       // packed target batch
  const uint32_t* __restrict__ q_lens;
  const uint32_t* __restrict__ t_lens;
  const uint32_t* __restrict__ q_offs;   // byte offsets (x8); word index = off >> 3
  const uint32_t* __restrict__ t_offs;
  const uint32_t* __restrict__ seed_scores;
  int n;                                 // pairs in this launch
  // scoring (positive penalties, src/gasal.h:160-165)
  int match, mismatch, gap_open, gap_extend;
  uint32_t n_value;                      // N_CODE & 15
  int n_score;                           // score of a pair involving N under the LOCAL rule (0 or -N_PENALTY)
  int n_is_wild_global;                  // 1 if N_PENALTY is defined (then the GLOBAL rule treats N specially too)
  // results (SoA, gasal_res_t arrays; any may be null)
  int32_t* score; int32_t* q_end; int32_t* t_end; int32_t* q_start; int32_t* t_start;
  int32_t* score2; int32_t* q_end2; int32_t* t_end2;
  // per-launch scratch: column state (H,E) as short2, laid out [column][pair] so a warp's accesses coalesce
  uint32_t* colstate; uint32_t col_stride; uint32_t col_rows;   // col_rows = columns available per pair
  // traceback
  uint4* tb; uint32_t tb_stride;         // tile t of pair p lives at tb[t * tb_stride + p]
  int tb_banded;                         // 1: only the tile band of banded.h:35,83-85 is stored (see band_geom)
  const uint64_t* tb_off;                // non-null (long pairs, wavefront kernel): pair p owns tiles tb[tb_off[p] + tile], sized by ITS OWN shape
  uint8_t* cigar; uint32_t* n_cigar_ops;
  int head, tail, k_band;
};

__device__ __forceinline__ uint32_t nib(uint32_t w, int k) { return (w >> (28 - 4 * k)) & 15u; }
__device__ __forceinline__ uint32_t n_words(uint32_t len) { return (len + 7u) >> 3; }

// reverse the order of the 8 nibbles of a word
__device__ __forceinline__ uint32_t nibrev(uint32_t x) {
  x = __byte_perm(x, 0, 0x0123);
  return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
}

// word j of the TRUE-LENGTH reversed sequence with zero pad nibbles -- what the reference builds in
// per-thread arrays before the semi-global start pass (semiglobal_kernel_template.h:232-266).
__device__ __forceinline__ uint32_t rev_true_word(const uint32_t* __restrict__ w, int len, int j) {
  const int hi = len - 1 - 8 * j;
  if (hi < 0) return 0u;
  const int lo = hi - 7;
  const int wi = lo >> 3;                 // arithmetic shift: floor
  const int sh = lo - 8 * wi;
  const uint32_t a = (wi >= 0) ? w[wi] : 0u;
  const uint32_t b = (sh != 0 || wi < 0) ? w[wi + 1] : 0u;
  const uint32_t x = __funnelshift_l(b, a, 4 * sh);
  return nibrev(x);
}

// 4-bit code -> 0..3 for A(1) C(3) G(7) T(4), 4 for anything else
__device__ __forceinline__ uint32_t base_idx(uint32_t code) {
  return (uint32_t)((0x4444444424431404ull >> (4 * code)) & 15ull);
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

// Tile band of one pair (banded.h:35,83-85): strip i covers query words [lo(i), hi(i)).  The banded traceback matrix
// keeps W = max width words per strip: tile of (strip i, word j, half) = (i * W + (j - lo(i))) * 2 + half.
struct BandGeom { int kw, k_other, W; };
__device__ __forceinline__ BandGeom band_geom(int QR, int TR, int k_band) {
  BandGeom b; b.kw = k_band >> 3; b.k_other = TR - (QR - b.kw);
  b.W = max(1, min(QR, b.kw + b.k_other - 1));
  return b;
}
__device__ __forceinline__ int band_lo(const BandGeom& b, int i) { return max(0, i - b.k_other + 1); }
__device__ __forceinline__ int band_hi(const BandGeom& b, int i, int QR) { return min(b.kw + i, QR); }

// ---- direction nibbles from packed s16x2 lanes (LOCAL WITH_TB; used by dp_s16.cuh and dp_wave.cuh; see dp_s16.cuh for the derivation)
__device__ __forceinline__ uint32_t lop3_code0(uint32_t a, uint32_t b, uint32_t s) {   // (~a & s) | (a & ~b)
  uint32_t d; asm("lop3.b32 %0, %1, %2, %3, 0x3a;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d;
}

__device__ __forceinline__ void tb_nibbles(uint32_t d, uint32_t dm, uint32_t h, uint32_t f, uint32_t e, uint32_t s2,
                                           uint32_t neg_ext2, uint32_t minus1_2, uint32_t mm, uint32_t sixteen, uint32_t& wL, uint32_t& wH) {
  const uint32_t fe = __vadd2(f, neg_ext2), ee = __vadd2(e, neg_ext2), dm1 = __vadd2(dm, minus1_2);
  const uint32_t qf = __vsub2(dm1, fe), qe = __vsub2(dm1, ee), a = __vsub2(d, h), b = __vsub2(f, h);
  const uint32_t c0 = lop3_code0(a, b, s2);
  const uint32_t g1 = prmt(c0, a, 0x7351);          // [c0.L, a.L, c0.H, a.H]   (sign-carrying bytes)
  const uint32_t g2 = prmt(qe, qf, 0x7351);         // [qe.L, qf.L, qe.H, qf.H]
  const uint32_t xl = prmt(g1, g2, 0x5410) & 0x80808080u, xh = prmt(g1, g2, 0x7632) & 0x80808080u;
  wL = wL * sixteen + ((xl * mm) >> 28);
  wH = wH * sixteen + ((xh * mm) >> 28);
}

__device__ __forceinline__ void tb_push(uint32_t (&d)[4], uint32_t w) { d[0] = d[1]; d[1] = d[2]; d[2] = d[3]; d[3] = w; }

__device__ __forceinline__ short2 u32_as_short2(uint32_t v) { short2 r; r.x = (short)(v & 0xFFFFu); r.y = (short)(v >> 16); return r; }
__device__ __forceinline__ uint32_t pack_short2(int h, int e) { return ((uint32_t)h & 0xFFFFu) | ((uint32_t)e << 16); }

}  // namespace gb200
