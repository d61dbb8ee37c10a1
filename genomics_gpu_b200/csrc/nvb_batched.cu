// nvb_batched.cu -- nvbio-style batched alignment SCORING on B200 (C ABI: include/nvbio_b200.h).
//
// What the reference does (Non-CDP/NvB/nvbio/alignment): BatchedAlignmentScore<stream, DeviceThreadScheduler>::enact
// (batched_inl.h) runs one CUDA thread per (pattern, text) job through aln::alignment_score (alignment.h:456-530), which
// for Gotoh is gotoh_alignment_score_dispatch::run (gotoh/gotoh_inl.h:473-900): bands of 8 pattern columns in int32
// registers, the band's right border (H, E) parked per text row in a short2 column in local / global memory.
//
// This module (same arithmetic, bit-exact -- oracle/nvbio_compat.c is the pinned restatement):
//   k_nvb_pack  one warp per sequence: bytes -> 2-bit codes, 16 per word, plus a per-sequence "not pure ACGT" flag
//   k_nvb_s16   fast path: ONE THREAD = TWO JOBS, one per 16-bit lane of the Blackwell DPX s16x2 instructions
//               (VIADDMNMX.S16x2, VIMNMX3.S16x2[.RELU]); bands of 16 pattern columns in registers; substitution scores
//               by one PRMT per cell pair from per-column byte tables (sign-replicating selector built once per text
//               row); H is kept as H + Go so that E, F and the diagonal each cost one instruction; the band border
//               goes through a coalesced [text row][thread] scratch (8 B per 32 cells).  Used when a host-side range
//               proof shows every DP value of the batch fits 16 bits; jobs containing a non-ACGT symbol are pushed to a
//               worklist and redone by
//   k_nvb_i32   general path: one thread per job, int32 DPX (VIADDMNMX, VIMNMX3[.RELU]), raw byte comparison (any
//               alphabet), asymmetric Smith-Waterman gap costs, and the BestSink coordinates in the reference's
//               report order.
// No CPU fallback: without a CUDA device every entry point fails with a message.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <limits.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/nvbio_b200.h"
#include "dp_launch.h"       // nvb_strip_launch: the quad-strip scoring kernel lives in the s16x2 translation unit (nvb_strip.cuh)

namespace nvb200 {

constexpr int kNeg32 = -(1 << 29);
constexpr int kNeg16 = -32000;
constexpr int kThreads = 128;
constexpr uint32_t kFull = 0xFFFFFFFFu;

struct NvbArgs {
  const uint8_t* pbytes; const uint32_t* poffs; const uint32_t* ppk; const uint8_t* pflag;
  const uint8_t* tbytes; const uint32_t* toffs; const uint32_t* tpk; const uint8_t* tflag;
  const uint32_t* pair_p; const uint32_t* pair_t; int text_shared;     // text_shared: every job uses text 0
  uint32_t first, count;                                               // job range of this launch
  int match, mismatch, go_t, ge_t, go_p, ge_p;                         // _t: gap consuming text (F), _p: consuming pattern (E)
  int band_shift;                                                      // log2 of the reference's BAND_LEN (tie order of LOCAL sinks)
  uint32_t sixteen;                                                    // == 16 (run-time operand: see k_nvb_s16)
  int32_t* score; int32_t* sink_i; int32_t* sink_j;
  uint2* scratch; uint32_t stride;                                     // band border (H, E) per text row: scratch[row * stride + thread] ...
  uint32_t rows_cap;                                                   // ... or, NVB_CTA_LAYOUT, scratch[(cta * rows_cap + row) * 128 + thread in cta]
  uint32_t* worklist; uint32_t* work_count;
  const int32_t* min_scores; int32_t min_score;                        // k_nvb_i32<.., MINSC = true>: per-job thresholds (or NULL: min_score for every job)
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t splat16(int v) { const uint32_t u = (uint32_t)v & 0xFFFFu; return u | (u << 16); }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(short)(v >> 16); }

// ------------------------------------------------------------------------------------------------------------------
// 2-bit packing: symbol k of a sequence in bits 2(k&15).. of word (k >> 4); sequence s starts at word (offs[s] >> 4) + s
__global__ void __launch_bounds__(256) k_nvb_pack(const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ offs, uint32_t nseq, int alphabet,
                                                  uint32_t* __restrict__ packed, uint8_t* __restrict__ flag) {
  const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nseq) return;
  const uint32_t o = offs[s], len = offs[s + 1] - o, wo = (o >> 4) + s, nw = (len + 15) >> 4;
  uint32_t bad = 0;
  for (uint32_t w = lane; w < nw; w += 32) {
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const uint32_t c = 16 * w + k;
      if (c < len) {
        const uint32_t b = bytes[o + c];
        uint32_t code;
        if (alphabet == NVBX_CODES) code = b < 4 ? b : 4;
        else code = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : 4;
        bad |= code >> 2;
        v |= (code & 3) << (2 * k);
      }
    }
    packed[wo + w] = v;
  }
  bad = __any_sync(kFull, bad != 0);
  if (lane == 0) flag[s] = (uint8_t)bad;
}

struct Job { uint32_t pi, ti, M, N; };
__device__ __forceinline__ Job load_job(const NvbArgs& a, uint32_t job) {
  Job j;
  j.pi = a.pair_p ? a.pair_p[job] : job;
  j.ti = a.text_shared ? 0u : (a.pair_t ? a.pair_t[job] : job);
  j.M = a.poffs[j.pi + 1] - a.poffs[j.pi];
  j.N = a.toffs[j.ti + 1] - a.toffs[j.ti];
  return j;
}

// Two formulations of the packed row update, chosen per alignment type from measurements (262144 x 150 x 600, TCUPS):
//   chain   : H kept as H + Go, E' = VIADDMNMX(E, Ge, Hg_left), H = VIMNMX3(E, F, D): 4 ALU + 2 FMA-pipe instructions per cell pair,
//             three dependent instructions per cell          semi-global 4.78, local 4.51, global 4.01
//   off-chain: E chain off H (see the kernel comment): 5 ALU + 1 FMA-pipe, one dependent instruction per cell
//                                                            semi-global 4.60, local 4.29, global 4.34
#ifndef NVB_CHAIN_MASK
#define NVB_CHAIN_MASK 0x6   // bit t set: alignment type t (0 GLOBAL, 1 LOCAL, 2 SEMI_GLOBAL) uses the chain formulation
#endif
// Band-border scratch layout: 1 = [CTA][text row][thread of the CTA] -- a CTA streams through its own contiguous region (1 KB per
// row, rows adjacent: DRAM-page friendly, compile-time row stride); 0 = [text row][every thread of the launch] (1 KB pieces 1 MB apart).
#ifndef NVB_CTA_LAYOUT
#define NVB_CTA_LAYOUT 1
#endif
#ifndef NVB_CHAIN_MASK_I32
#define NVB_CHAIN_MASK_I32 0x0  // same choice for the int32 kernel k_nvb_i32w (measured: no gain, 2.15 either way; global 2.01 with the chain)
#endif
__host__ __device__ constexpr bool chain_for(int type) { return (NVB_CHAIN_MASK >> type) & 1; }
#ifndef NVB_ROW_UNROLL
#define NVB_ROW_UNROLL 4
#endif
#define NVB_PRAGMA(x) _Pragma(#x)
#define NVB_UNROLL(n) NVB_PRAGMA(unroll n)      // nvcc wants a literal unroll factor
#ifndef NVB_PREFETCH_ROWS
#define NVB_PREFETCH_ROWS 7
#endif
constexpr uint32_t kFixedStride = 1u << 17;   // threads (job pairs) per fast-path launch; 7 rows x 1 MB stays inside the 24-bit load offset
constexpr int kPrefetchRows = NVB_PREFETCH_ROWS;
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// ------------------------------------------------------------------------------------------------------------------
// general path: one thread per job, int32, bands of 8 pattern columns (gotoh_inl.h:473-593 with BAND_LEN = 8)
template <int N> __device__ __forceinline__ int select_reg(const int (&v)[N], int k) {
  int x = v[N - 1];
#pragma unroll
  for (int j = 0; j < N - 1; j++) x = (k == j) ? v[j] : x;
  return x;
}
// is cell (i, j) reported after cell (bi, bj) in the reference's order (band of columns, then row, then column)?
__device__ __forceinline__ bool reported_later(int i, int j, int bi, int bj, int band_shift) {
  if (bj < 0) return true;
  const int b1 = (j - 1) >> band_shift, b0 = (bj - 1) >> band_shift;
  if (b1 != b0) return b1 > b0;
  if (i != bi) return i > bi;
  return j > bj;
}

// MINSC: alignment_score's min_score threshold (alignment.h:456-530).  After each of the reference's bands of BAND_LEN = 1 <<
// band_shift pattern columns except the last one, max_score = the maximum over the text rows of H(i, last column of the band); the
// job gives up when max_score + (columns still to come) * match < min_score (gotoh_inl.h:826-830, sw_inl.h:699-701).  LOCAL keeps
// what the bands processed so far reported; GLOBAL / SEMI_GLOBAL report from the last band only, so they return BestSink's initial
// state: score -(1 << 30) (Field_traits<int32>::min()), sink (-1, -1).  This kernel's bands of 8 ARE the reference's for Gotoh;
// for Smith-Waterman / edit distance (bands of 16) the check falls on every second band.
template <int TYPE, bool SINK, bool MINSC = false>
__global__ void __launch_bounds__(kThreads) k_nvb_i32(const NvbArgs a, int from_worklist) {
  const uint32_t idx = blockIdx.x * kThreads + threadIdx.x;
  uint32_t job;
  if (from_worklist) { if (idx >= *a.work_count) return; job = a.worklist[idx]; }
  else { if (idx >= a.count) return; job = a.first + idx; }
  const Job jb = load_job(a, job);
  const uint8_t* __restrict__ P = a.pbytes + a.poffs[jb.pi];
  const uint8_t* __restrict__ T = a.tbytes + a.toffs[jb.ti];
  const int M = (int)jb.M, N = (int)jb.N;
  uint2* __restrict__ col = NVB_CTA_LAYOUT ? a.scratch + (size_t)blockIdx.x * a.rows_cap * kThreads + threadIdx.x : a.scratch + idx;
  const uint32_t stride = NVB_CTA_LAYOUT ? (uint32_t)kThreads : a.stride;
  const int go_t = a.go_t, ge_t = a.ge_t, go_p = a.go_p, ge_p = a.ge_p, match = a.match, mismatch = a.mismatch;

  int best = INT_MIN, bi = -1, bj = -1;
  const int min_score = MINSC ? (a.min_scores ? a.min_scores[job] : a.min_score) : 0;
  const int ref_bands = (M + (1 << a.band_shift) - 1) >> a.band_shift;
  for (int block = 0; block < M; block += 8) {
    int colmax = INT_MIN;                                                                   // MINSC: max over the rows of H(i, block + 8)
    int q[8];
#pragma unroll
    for (int j = 0; j < 8; j++) q[j] = (block + j < M) ? (int)P[block + j] : 256 + j;      // pad columns never match
    int H[9], F[9];
#pragma unroll
    for (int j = 0; j <= 8; j++) {
      const int c = block + j;
      H[j] = (TYPE == NVBX_LOCAL || c == 0) ? 0 : go_p + ge_p * (c - 1);                    // gotoh_inl.h:690-695
      F[j] = kNeg32;
    }
    const bool first = block == 0, last = block + 8 >= M;
    const int mj = last ? M - block : 8;                                                     // valid columns of this band
    uint2 nxt = first ? make_uint2(0, 0) : col[0];
    for (int i = 1; i <= N; i++) {
      int hl, e;
      if (first) { hl = (TYPE == NVBX_GLOBAL) ? go_t + ge_t * (i - 1) : 0; e = (TYPE == NVBX_LOCAL) ? 0 : kNeg32; }   // gotoh_inl.h:77-88
      else { hl = (int)nxt.x; e = (int)nxt.y; if (i < N) nxt = col[(size_t)i * stride]; }
      const int t = (int)T[i - 1];
      int diag = H[0];
      H[0] = hl;
#pragma unroll
      for (int j = 1; j <= 8; j++) {
        F[j] = __viaddmax_s32(F[j], ge_t, H[j] + go_t);
        e = __viaddmax_s32(e, ge_p, H[j - 1] + go_p);
        const int s = (t == q[j - 1]) ? match : mismatch;
        const int d = diag + s;
        const int h = (TYPE == NVBX_LOCAL) ? __vimax3_s32_relu(e, F[j], d) : __vimax3_s32(e, F[j], d);
        diag = H[j];
        H[j] = h;
      }
      if (!last) col[(size_t)(i - 1) * stride] = make_uint2((uint32_t)H[8], (uint32_t)e);
      if (MINSC) colmax = max(colmax, H[8]);
      if (TYPE == NVBX_LOCAL) {
        int rowmax = INT_MIN;
#pragma unroll
        for (int j = 1; j <= 8; j++) if (j <= mj) rowmax = max(rowmax, H[j]);
        if (SINK) {
          if (rowmax >= best) {
            int cj = 0;
#pragma unroll
            for (int j = 1; j <= 8; j++) if (j <= mj && H[j] == rowmax) cj = j;               // largest column of the row
            if (rowmax > best || reported_later(i, block + cj, bi, bj, a.band_shift)) { best = rowmax; bi = i; bj = block + cj; }
          }
        } else best = max(best, rowmax);
      } else if (last) {
        const int hM = select_reg(H, mj);
        if (TYPE == NVBX_SEMI_GLOBAL) { if (hM >= best) { best = hM; bi = i; bj = M; } }    // utils_inl.h:279-299, BestSink `<=`
        else if (i == N) { best = hM; bi = N; bj = M; }                                      // gotoh_inl.h:896-897
      }
    }
    if (MINSC) {
      const int done = block + 8;                                                            // pattern columns processed so far
      if ((done & ((1 << a.band_shift) - 1)) == 0 && (done >> a.band_shift) < ref_bands &&
          (long long)colmax + (long long)(M - done) * match < (long long)min_score) {
        if (TYPE != NVBX_LOCAL) { best = -(1 << 30); bi = -1; bj = -1; }
        break;
      }
    }
  }
  a.score[job] = best;
  if (SINK) { a.sink_i[job] = bi; a.sink_j[job] = bj; }
}

// ------------------------------------------------------------------------------------------------------------------
// int32 path for ordered gap costs (Go <= Ge <= 0 on both sides): one thread per job, bands of 16 pattern columns, the same
// formulation as the s16x2 kernel below (H as is, F - Go_t, E - Go_p; E chain off H; X written in place right-to-left;
// GLOBAL / SEMI_GLOBAL patterns right-aligned so H(i, M) is register H[16]).  Adds what the packed kernel cannot do: raw byte
// comparison (any alphabet), asymmetric Smith-Waterman costs, values beyond 16 bits, and the LOCAL BestSink coordinates: the
// winner is the maximum by (score, report order key) with key = (band of the reference's BAND_LEN columns, text row,
// column) -- compared explicitly, so it does not depend on this kernel's own traversal order.
constexpr int kFrontGlobal = -(1 << 20);        // substitution score of GLOBAL's virtual front columns: the diagonal never wins

template <int TYPE, bool SINK>
__global__ void __launch_bounds__(kThreads, 4) k_nvb_i32w(const NvbArgs a, int from_worklist) {
  const uint32_t idx = blockIdx.x * kThreads + threadIdx.x;
  // worklist mode: the grid is bounded by the scratch (one band-border column per THREAD), so a thread may take several entries
  const uint32_t limit = from_worklist ? *a.work_count : a.count, step = gridDim.x * kThreads;
  for (uint32_t widx = idx; widx < limit; widx += step) {
  const uint32_t job = from_worklist ? a.worklist[widx] : a.first + widx;
  const Job jb = load_job(a, job);
  const uint8_t* __restrict__ P = a.pbytes + a.poffs[jb.pi];
  const uint8_t* __restrict__ T = a.tbytes + a.toffs[jb.ti];
  const int M = (int)jb.M, N = (int)jb.N;
  const int d = (TYPE == NVBX_LOCAL) ? 0 : (16 - (M & 15)) & 15;                           // virtual columns in front of the pattern
  const int nb = (M + d + 15) >> 4;
  uint2* __restrict__ col = NVB_CTA_LAYOUT ? a.scratch + (size_t)blockIdx.x * a.rows_cap * kThreads + threadIdx.x : a.scratch + idx;
  const size_t stride = NVB_CTA_LAYOUT ? (size_t)kThreads : (size_t)a.stride;
  const int go_t = a.go_t, ge_t = a.ge_t, go_p = a.go_p, ge_p = a.ge_p, match = a.match, mismatch = a.mismatch;
  const int front = (TYPE == NVBX_GLOBAL) ? kFrontGlobal : 0;
  constexpr bool CHAIN = (NVB_CHAIN_MASK_I32 >> TYPE) & 1;                                  // chain: H kept as H + Go_t (see k_nvb_s16)
  const int tsh = CHAIN ? go_t : 0, dgo = go_p - go_t, matchs = match - tsh;

  int best = INT_MIN, bi = -1, bj = -1;
  for (int b = 0; b < nb; b++) {
    int q[16], sx[16];                                                                    // symbol (256+: never equal) and non-match score per column
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int c = 16 * b + j;
      const bool real = c >= d && c < M + d;
      q[j] = real ? (int)P[c - d] : 256 + j;
      sx[j] = ((c < d) ? front : mismatch) - tsh;
    }
    int H[17], F[17];                                                                     // F holds F - Go_t
#pragma unroll
    for (int j = 0; j <= 16; j++) {
      const int jr = 16 * b + j - d;                                                       // real 1-based column (<= 0: column 0)
      H[j] = ((TYPE == NVBX_LOCAL || jr <= 0) ? 0 : go_p + ge_p * (jr - 1)) + tsh;
      F[j] = kNeg32;
    }
    const bool first = b == 0, last = b == nb - 1;
    uint2* rp = col;
    uint2 nxt = first ? make_uint2(0, 0) : col[0];
    if (!first) {
#pragma unroll
      for (int r = 1; r < kPrefetchRows; r++) if (r < N) prefetch_l1(col + (size_t)r * stride);
    }
NVB_UNROLL(2)
    for (int i = 1; i <= N; i++) {
      int hl, e;                                                                           // e holds E - Go_p
      if (first) {
        hl = ((TYPE == NVBX_GLOBAL) ? go_t + ge_t * (i - 1) : 0) + tsh;
        e = (TYPE == NVBX_LOCAL) ? (CHAIN ? 0 : -go_p) : kNeg32;
      } else {
        hl = (int)nxt.x; e = (int)nxt.y;
        if (i < N) nxt = rp[stride];
        if (i + kPrefetchRows <= N) prefetch_l1(rp + (size_t)kPrefetchRows * stride);
      }
      const int t = (int)T[i - 1];
      if (CHAIN) {
        // H holds H + Go_t, e the true E: F, then D in place (right-to-left); E, H = max(E, F, D), H + Go_t (left-to-right)
#pragma unroll
        for (int j = 16; j >= 1; j--) {
          F[j] = __viaddmax_s32(F[j], ge_t, H[j]);
          H[j] = H[j - 1] + ((t == q[j - 1]) ? matchs : sx[j - 1]);
        }
        H[0] = hl;
#pragma unroll
        for (int j = 1; j <= 16; j++) {
          e = __viaddmax_s32(e, ge_p, H[j - 1] + dgo);
          const int h = (TYPE == NVBX_LOCAL) ? __vimax3_s32_relu(e, F[j], H[j]) : __vimax3_s32(e, F[j], H[j]);
          H[j] = h + tsh;
        }
      } else {
#pragma unroll
      for (int j = 16; j >= 1; j--) {
        F[j] = __viaddmax_s32(F[j], ge_t, H[j]);
        const int dd = H[j - 1] + ((t == q[j - 1]) ? match : sx[j - 1]);
        H[j] = __viaddmax_s32(F[j], go_t, dd);                                               // X = max(F, D)
      }
      H[0] = hl;
      int eprev = e;
#pragma unroll
      for (int j = 1; j <= 16; j++) {
        e = __viaddmax_s32(e, ge_p, H[j - 1]);
        if (j > 1) H[j - 1] = (TYPE == NVBX_LOCAL) ? __viaddmax_s32_relu(eprev, go_p, H[j - 1]) : __viaddmax_s32(eprev, go_p, H[j - 1]);
        eprev = e;
      }
      H[16] = (TYPE == NVBX_LOCAL) ? __viaddmax_s32_relu(e, go_p, H[16]) : __viaddmax_s32(e, go_p, H[16]);
      }
      if (!last) *rp = make_uint2((uint32_t)H[16], (uint32_t)e);
      rp += stride;
      if (TYPE == NVBX_LOCAL) {
        const int mj = last ? M - 16 * b : 16;                                               // valid columns of this band (pad columns never win, but keep exactness)
        int rowmax = H[1];
#pragma unroll
        for (int j = 2; j <= 16; j += 2) rowmax = (j < 16) ? __vimax3_s32(rowmax, H[j], H[j + 1]) : max(rowmax, H[j]);
        if (SINK) {
          if (rowmax >= best) {
            int cj = 0;
#pragma unroll
            for (int j = 1; j <= 16; j++) if (j <= mj && H[j] == rowmax) cj = j;             // largest column of the row
            if (cj > 0 && (rowmax > best || reported_later(i, 16 * b + cj, bi, bj, a.band_shift))) { best = rowmax; bi = i; bj = 16 * b + cj; }
          }
        } else best = max(best, rowmax);
      } else if (last) {
        if (TYPE == NVBX_SEMI_GLOBAL) { if (H[16] >= best) { best = H[16]; bi = i; bj = M; } }
        else if (i == N) { best = H[16]; bi = N; bj = M; }
      }
    }
  }
  a.score[job] = best - tsh;                                                               // tracked in the kernel's H frame
  if (SINK) { a.sink_i[job] = bi; a.sink_j[job] = bj; }
  }   // worklist / job loop
}

// ------------------------------------------------------------------------------------------------------------------
// fast path: one thread = two jobs (lane L = job 2k, lane H = job 2k+1), s16x2, bands of 16 pattern columns, score only.
//
// Frame and recurrence (symmetric gap costs with Go <= Ge <= 0, proven on the host): H is kept as it is, F and E are kept as
// Fg = F - Go and Eg = E - Go.  Because Go <= Ge, E(i,j+1) = max(E(i,j) + Ge, H(i,j) + Go) = max(E(i,j) + Ge, X(i,j) + Go) with
// X = max(F, D), which takes H off the in-row dependency chain (the chain is ONE VIADDMNMX per cell; all X of a row are
// independent).  Per cell pair:
//     Fg = VIADDMNMX(Fg, Ge, H_top)                 ALU     = max(F + Ge, H_top + Go) - Go
//     s  = PRMT(tabA[j], tabB[j], row selector)     ALU     substitution scores of both jobs
//     D  = VIADD.16x2(H_diag, s)                    FMA pipe
//     X  = VIADDMNMX(Fg, Go, D)                     ALU     written in place over H[j] (right-to-left pass: no rotating copy)
//     Eg'= VIADDMNMX(Eg, Ge, X_left)                ALU     (left-to-right pass)
//     H  = VIADDMNMX[.RELU](Eg, Go, X)              ALU     = max(E, F, D [, 0])
// GLOBAL / SEMI_GLOBAL read H(i, M): each job's pattern is RIGHT-aligned in its bands (d = -M mod 16 virtual columns in front),
// so column M is always register H[16] of the job's last band -- no register select chain.  The virtual columns reproduce
// column 0 by construction: substitution 0 (SEMI_GLOBAL: H stays 0) or -128 (GLOBAL: H = Go + Ge (i-1) through F).
__device__ __forceinline__ uint32_t band_codes(const uint32_t* __restrict__ pk, int b, int d, int nw) {
  const uint32_t hi = (b < nw) ? pk[b] : 0u, lo = (b >= 1 && b - 1 < nw) ? pk[b - 1] : 0u;
  return d ? __funnelshift_r(lo, hi, 32 - 2 * d) : hi;
}

// LOCAL sinks on the packed path.  Per cell key = H * 16 + (column - 1 within the band); the row's largest key is its candidate
// (largest column among the row's maxima).  The reference's report order is (band of BAND_LEN columns, text row, column) with the
// LAST maximum winning: for BAND_LEN = 8 a band of 16 holds two reference bands, so rows are compared on key >> 3 = (H, half) and
// a later row wins ties; for BAND_LEN = 16 on key >> 4 = H.  best = max(best, c) per lane; lanes where c >= best record the row's
// key and position.  One asm block so that ptxas fuses the lane compares into the predicate outputs of VIMNMX.S16x2.
struct SinkTrack { uint32_t bestc, keyL, keyH, posL, posH; };
__device__ __forceinline__ void sink_update(SinkTrack& t, uint32_t c, uint32_t rowkey, uint32_t pos) {
  asm volatile(
      "{\n\t"
      ".reg .pred pl, ph;\n\t"
      ".reg .s16 n0, n1, c0, c1;\n\t"
      ".reg .b32 nb;\n\t"
      "max.s16x2 nb, %0, %5;\n\t"
      "mov.b32 {n0, n1}, nb;\n\t"
      "mov.b32 {c0, c1}, %5;\n\t"
      "setp.eq.s16 pl, n0, c0;\n\t"
      "setp.eq.s16 ph, n1, c1;\n\t"
      "mov.b32 %0, nb;\n\t"
      "@pl mov.b32 %1, %6;\n\t"
      "@pl mov.b32 %3, %7;\n\t"
      "@ph mov.b32 %2, %6;\n\t"
      "@ph mov.b32 %4, %7;\n\t"
      "}"
      : "+r"(t.bestc), "+r"(t.keyL), "+r"(t.keyH), "+r"(t.posL), "+r"(t.posH)
      : "r"(c), "r"(rowkey), "r"(pos));
}

// FIXED: the scratch row stride is the compile-time kFixedStride, so the next-row load, the prefetch and the store share ONE
// pointer with immediate offsets (the run-time-stride variant serves batches whose texts are too long for that layout).
template <int TYPE, bool FIXED, bool SINK = false>
__global__ void __launch_bounds__(kThreads, 4) k_nvb_s16(const NvbArgs a) {
  const uint32_t idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= (a.count + 1) / 2) return;
  const uint32_t jobA = a.first + 2 * idx, jobB = min(jobA + 1, a.first + a.count - 1);
  const Job A = load_job(a, jobA), B = load_job(a, jobB);
  if (a.pflag[A.pi] | a.tflag[A.ti] | a.pflag[B.pi] | a.tflag[B.ti]) {                       // a symbol outside ACGT: general kernel
    const uint32_t n = (jobB != jobA) ? 2u : 1u;
    const uint32_t at = atomicAdd(a.work_count, n);
    a.worklist[at] = jobA;
    if (n == 2) a.worklist[at + 1] = jobB;
    return;
  }
  const uint32_t* __restrict__ pA = a.ppk + (a.poffs[A.pi] >> 4) + A.pi;
  const uint32_t* __restrict__ pB = a.ppk + (a.poffs[B.pi] >> 4) + B.pi;
  const uint32_t* __restrict__ tA = a.tpk + (a.toffs[A.ti] >> 4) + A.ti;
  const uint32_t* __restrict__ tB = a.tpk + (a.toffs[B.ti] >> 4) + B.ti;
  const int MA = (int)A.M, MB = (int)B.M, NA = (int)A.N, NB = (int)B.N;
  const int Nx = max(NA, NB);
  const int dA = (TYPE == NVBX_LOCAL) ? 0 : (16 - (MA & 15)) & 15, dB = (TYPE == NVBX_LOCAL) ? 0 : (16 - (MB & 15)) & 15;
  const int nbA = (MA + dA + 15) >> 4, nbB = (MB + dB + 15) >> 4, nb = max(nbA, nbB);
  const int nwA = (MA + 15) >> 4, nwB = (MB + 15) >> 4;
  uint2* __restrict__ col = NVB_CTA_LAYOUT ? a.scratch + (size_t)blockIdx.x * a.rows_cap * kThreads + threadIdx.x : a.scratch + idx;
  const size_t stride = NVB_CTA_LAYOUT ? (size_t)kThreads : FIXED ? (size_t)kFixedStride : (size_t)a.stride;
  const int go = a.go_t, ge = a.ge_t;
  const uint32_t go2 = splat16(go), ge2 = splat16(ge);
  constexpr bool CHAIN = chain_for(TYPE);
  const int tsh = CHAIN ? go : 0;                                                          // chain: H is kept as H + Go, the tables hold s - Go
  const uint32_t mis4 = ((uint32_t)(a.mismatch - tsh) & 0xFFu) * 0x01010101u;
  const uint32_t delta = ((uint32_t)(a.match - tsh) ^ (uint32_t)(a.mismatch - tsh)) & 0xFFu;
  const uint32_t front4 = (TYPE == NVBX_GLOBAL) ? 0x80808080u : ((uint32_t)(0 - tsh) & 0xFFu) * 0x01010101u;   // virtual columns in front of the pattern

  int bestA = INT_MIN, bestB = INT_MIN, rowA = -1, rowB = -1;                              // rowA/rowB: SEMI_GLOBAL sinks (last maximum wins)
  uint32_t best2 = (TYPE == NVBX_LOCAL) ? 0u : 0x80008000u;
  SinkTrack trk = {0u, 0u, 0u, 0u, 0u};                                                    // LOCAL + SINK
  const uint32_t sixteen = a.sixteen, kshift = (uint32_t)a.band_shift;                       // key >> 3 (BAND_LEN 8) or >> 4 (16); 16 at run time keeps the key an IMAD
  const uint32_t kmask = (0xFFFFu >> kshift) * 0x00010001u;
  for (int b = 0; b < nb; b++) {
    const uint32_t wA = band_codes(pA, b, dA, nwA), wB = band_codes(pB, b, dB, nwB);
    uint32_t tabA[16], tabB[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int c = 16 * b + j;
      tabA[j] = (c < dA) ? front4 : (c < MA + dA) ? (mis4 ^ (delta << (8 * ((wA >> (2 * j)) & 3u)))) : mis4;
      tabB[j] = (c < dB) ? front4 : (c < MB + dB) ? (mis4 ^ (delta << (8 * ((wB >> (2 * j)) & 3u)))) : mis4;
    }
    uint32_t H[17], F[17];                                                               // F holds Fg = F - Go
#pragma unroll
    for (int j = 0; j <= 16; j++) {
      const int ja = 16 * b + j - dA, jb = 16 * b + j - dB;                                // real 1-based column (<= 0: column 0)
      const int va = (TYPE == NVBX_LOCAL || ja <= 0) ? 0 : go + ge * (ja - 1), vb = (TYPE == NVBX_LOCAL || jb <= 0) ? 0 : go + ge * (jb - 1);
      H[j] = ((uint32_t)(va + tsh) & 0xFFFFu) | ((uint32_t)(vb + tsh) << 16);
      F[j] = splat16(kNeg16);
    }
    const bool first = b == 0, last = b == nb - 1;
    const bool capA = b == nbA - 1, capB = b == nbB - 1;
    if (TYPE == NVBX_LOCAL && SINK && a.band_shift == 3) trk.bestc &= 0xFFFEFFFEu;          // a later band wins ties against either half of an earlier one
    // band border of the previous band: read one row ahead into registers, kPrefetchRows rows ahead into L1 (the scratch of the
    // resident threads is larger than L2: without the prefetch the kernel waits on these loads -- ncu long-scoreboard stalls)
    uint2* rp = col;                                                                       // row i of the band border; row i+1 is loaded, row i+kPrefetchRows prefetched
    uint2 nxt = first ? make_uint2(0, 0) : col[0];
    if (!first) {
#pragma unroll
      for (int r = 1; r < kPrefetchRows; r++) if (r < Nx) prefetch_l1(col + (size_t)r * stride);
    }
    uint32_t twA = 0, twB = 0;
NVB_UNROLL(NVB_ROW_UNROLL)
    for (int i = 1; i <= Nx; i++) {
      if (((i - 1) & 15) == 0) { twA = (i <= NA) ? tA[(i - 1) >> 4] : 0u; twB = (i <= NB) ? tB[(i - 1) >> 4] : 0u; }
      const uint32_t sel = (twA & 3u) * 0x0011u + (twB & 3u) * 0x1100u + 0xC480u;            // bytes: [tA], sign, [4 + tB], sign
      twA >>= 2; twB >>= 2;
      uint32_t hl, e;                                                                      // e holds Eg = E - Go
      if (first) {
        hl = splat16(((TYPE == NVBX_GLOBAL) ? go + ge * (i - 1) : 0) + tsh);
        e = (TYPE == NVBX_LOCAL) ? splat16(CHAIN ? 0 : -go) : splat16(kNeg16);
      } else {
        hl = nxt.x; e = nxt.y;
        if (i < Nx) nxt = rp[stride];
        if (i + kPrefetchRows <= Nx) prefetch_l1(rp + (size_t)kPrefetchRows * stride);
      }
      uint32_t mrow = 0;                                                                   // LOCAL: row maximum of H (score) or of the keys (SINK)
      if (CHAIN) {
      // right-to-left: F and the diagonal sum D (in place over H[j]: H[j-1] still holds the previous row)
#pragma unroll
      for (int j = 16; j >= 1; j--) {
        F[j] = __viaddmax_s16x2(F[j], ge2, H[j]);
        H[j] = __vadd2(H[j - 1], prmt(tabA[j - 1], tabB[j - 1], sel));
      }
      H[0] = hl;
      // left-to-right: E, H = max(E, F, D), H + Go
      uint32_t hodd = 0;
#pragma unroll
      for (int j = 1; j <= 16; j++) {
        e = __viaddmax_s16x2(e, ge2, H[j - 1]);
        const uint32_t h = (TYPE == NVBX_LOCAL) ? __vimax3_s16x2_relu(e, F[j], H[j]) : __vimax3_s16x2(e, F[j], H[j]);
        H[j] = __vadd2(h, go2);
        if (TYPE == NVBX_LOCAL) {
          const uint32_t v = SINK ? h * sixteen + (uint32_t)(j - 1) * 0x00010001u : h;
          if (j & 1) hodd = v; else mrow = __vimax3_s16x2(mrow, hodd, v);
        }
      }
      } else {
      // right-to-left: F, diagonal, X (in place: H[j-1] still holds the previous row when H[j] is overwritten)
#pragma unroll
      for (int j = 16; j >= 1; j--) {
        F[j] = __viaddmax_s16x2(F[j], ge2, H[j]);
        const uint32_t d = __vadd2(H[j - 1], prmt(tabA[j - 1], tabB[j - 1], sel));
        H[j] = __viaddmax_s16x2(F[j], go2, d);
      }
      H[0] = hl;
      // left-to-right: the E chain (one instruction per cell) and the final H
      uint32_t eprev = e;
#pragma unroll
      for (int j = 1; j <= 16; j++) {
        e = __viaddmax_s16x2(e, ge2, H[j - 1]);                                              // Eg(i,j): H[j-1] is X(i,j-1) or the full H (same result)
        if (j > 1) H[j - 1] = (TYPE == NVBX_LOCAL) ? __viaddmax_s16x2_relu(eprev, go2, H[j - 1]) : __viaddmax_s16x2(eprev, go2, H[j - 1]);
        eprev = e;
      }
      H[16] = (TYPE == NVBX_LOCAL) ? __viaddmax_s16x2_relu(e, go2, H[16]) : __viaddmax_s16x2(e, go2, H[16]);
      if (TYPE == NVBX_LOCAL) {
#pragma unroll
        for (int j = 1; j <= 16; j += 2)
          mrow = SINK ? __vimax3_s16x2(mrow, H[j] * sixteen + (uint32_t)(j - 1) * 0x00010001u, H[j + 1] * sixteen + (uint32_t)j * 0x00010001u)
                      : __vimax3_s16x2(mrow, H[j], H[j + 1]);
      }
      }
      if (!last) *rp = make_uint2(H[16], e);
      rp += stride;
      if (TYPE == NVBX_LOCAL) {
        const uint32_t mask = (i <= NA ? 0x0000FFFFu : 0u) | (i <= NB ? 0xFFFF0000u : 0u);   // rows past a job's text end do not count
        if (SINK) {
          const uint32_t rk = mrow & mask;
          sink_update(trk, (rk >> kshift) & kmask, rk, ((uint32_t)b << 20) | (uint32_t)i);
        } else {
          const uint32_t m = __vmaxs2(best2, mrow);
          best2 = (m & mask) | (best2 & ~mask);
        }
      } else if (TYPE == NVBX_SEMI_GLOBAL) {
        if (SINK) {                                                                        // BestSink `score <= new` (sink_inl.h:52-60): the last row wins
          if (capA && i <= NA) { const int h = lo16(H[16]); if (h >= bestA) { bestA = h; rowA = i; } }
          if (capB && i <= NB) { const int h = hi16(H[16]); if (h >= bestB) { bestB = h; rowB = i; } }
        } else if (capA | capB) {
          const uint32_t mask = ((capA && i <= NA) ? 0x0000FFFFu : 0u) | ((capB && i <= NB) ? 0xFFFF0000u : 0u);
          best2 = __vmaxs2(best2, (H[16] & mask) | (0x80008000u & ~mask));
        }
      } else {
        if (capA && i == NA) bestA = lo16(H[16]);
        if (capB && i == NB) bestB = hi16(H[16]);
      }
    }
  }
  if (TYPE == NVBX_LOCAL || (TYPE == NVBX_SEMI_GLOBAL && !SINK)) { bestA = lo16(best2); bestB = hi16(best2); }
  if (TYPE != NVBX_LOCAL) { bestA -= tsh; bestB -= tsh; }                                  // captured H[16] carries the frame shift
  a.score[jobA] = bestA;
  if (jobB != jobA) a.score[jobB] = bestB;
  if (SINK && TYPE == NVBX_LOCAL) {
    // a score of 0 means every cell tied at 0: the last reported cell (N, M) wins (sink_inl.h:52-60)
    const uint32_t kA = trk.keyL & 0xFFFFu, kB = trk.keyH >> 16;
    const int sA = (int)(kA >> 4), sB = (int)(kB >> 4);
    a.score[jobA] = sA;
    a.sink_i[jobA] = sA ? (int)(trk.posL & 0xFFFFFu) : NA; a.sink_j[jobA] = sA ? (int)(16 * (trk.posL >> 20) + (kA & 15u) + 1) : MA;
    if (jobB != jobA) {
      a.score[jobB] = sB;
      a.sink_i[jobB] = sB ? (int)(trk.posH & 0xFFFFFu) : NB; a.sink_j[jobB] = sB ? (int)(16 * (trk.posH >> 20) + (kB & 15u) + 1) : MB;
    }
  } else if (SINK) {                                                                       // GLOBAL: (N, M); SEMI_GLOBAL: (last best row, M)
    a.sink_i[jobA] = (TYPE == NVBX_GLOBAL) ? NA : rowA; a.sink_j[jobA] = MA;
    if (jobB != jobA) { a.sink_i[jobB] = (TYPE == NVBX_GLOBAL) ? NB : rowB; a.sink_j[jobB] = MB; }
  }
}

}  // namespace nvb200
#include "nvb_tb.cuh"       // banded scoring + traceback kernels (k_nvb_banded, k_nvb_tb_fwd, k_nvb_tb_walk, k_nvb_banded_walk)
namespace nvb200 {

// ------------------------------------------------------------------------------------------------------------------
thread_local std::string g_err;
uint64_t g_launches = 0;
int g_policy = 0;

static int fail(const std::string& m) { g_err = m; fprintf(stderr, "[NVBX ERROR:] %s\n", m.c_str()); return -1; }
#define NVB_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(std::string(#x) + ": " + cudaGetErrorString(e_)); return -1; } } while (0)
#define NVB_CHECKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(std::string(#x) + ": " + cudaGetErrorString(e_)); nvbx_stream_destroy(s); return nullptr; } } while (0)

}  // namespace nvb200

struct nvbx_stream {
  cudaStream_t st = nullptr;
  uint8_t *pbytes = nullptr, *tbytes = nullptr, *pflag = nullptr, *tflag = nullptr;
  uint32_t *poffs = nullptr, *toffs = nullptr, *ppk = nullptr, *tpk = nullptr, *pair_p = nullptr, *pair_t = nullptr;
  uint32_t n_patterns = 0, n_texts = 0, n_pairs = 0, max_m = 0, max_n = 0;
  int text_shared = 0;
  uint64_t cells = 0;
  int32_t *score = nullptr, *sink_i = nullptr, *sink_j = nullptr;
  uint2* scratch = nullptr; size_t scratch_elems = 0;
  uint32_t *worklist = nullptr, *work_count = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  uint32_t* dirs = nullptr; size_t dirs_words = 0;                 // flow words of one chunk of jobs (traceback)
  int32_t *src_i = nullptr, *src_j = nullptr; uint32_t* n_ops = nullptr; uint8_t* ops = nullptr; uint32_t ops_stride = 0;
  bool have_tb = false;
  bool have_sinks = false, enacted = false;
  int32_t* min_scores = nullptr; int32_t min_score = INT32_MIN; bool use_min_score = false;     // nvbx_stream_set_min_scores
  std::string kernel = "none";
};

using namespace nvb200;

extern "C" {

void nvbx_set_policy(int policy) { g_policy = policy; }
uint64_t nvbx_kernel_launches(void) { return g_launches; }
const char* nvbx_last_error(void) { return g_err.c_str(); }

void nvbx_stream_destroy(nvbx_stream_t* s) {
  if (!s) return;
  cudaFree(s->pbytes); cudaFree(s->tbytes); cudaFree(s->pflag); cudaFree(s->tflag); cudaFree(s->poffs); cudaFree(s->toffs);
  cudaFree(s->ppk); cudaFree(s->tpk); cudaFree(s->pair_p); cudaFree(s->pair_t); cudaFree(s->score); cudaFree(s->sink_i); cudaFree(s->sink_j);
  cudaFree(s->scratch); cudaFree(s->worklist); cudaFree(s->work_count);
  cudaFree(s->min_scores); cudaFree(s->dirs); cudaFree(s->src_i); cudaFree(s->src_j); cudaFree(s->n_ops); cudaFree(s->ops);
  if (s->e0) cudaEventDestroy(s->e0);
  if (s->e1) cudaEventDestroy(s->e1);
  delete s;
}

nvbx_stream_t* nvbx_stream_create(const nvbx_batch_t* b, void* cuda_stream) {
  if (!b || !b->patterns || !b->texts || !b->pattern_offsets || !b->text_offsets || b->n_patterns == 0 || b->n_texts == 0 || b->n_pairs == 0) {
    fail("nvbx_stream_create: empty or NULL batch"); return nullptr;
  }
  if (b->alphabet != NVBX_CODES && b->alphabet != NVBX_ASCII) { fail("nvbx_stream_create: unknown alphabet"); return nullptr; }
  if (!b->pair_pattern && b->n_pairs > b->n_patterns) { fail("nvbx_stream_create: n_pairs > n_patterns without pair_pattern"); return nullptr; }
  const int text_shared = (!b->pair_text && b->n_texts == 1) ? 1 : 0;
  if (!b->pair_text && !text_shared && b->n_pairs > b->n_texts) { fail("nvbx_stream_create: n_pairs > n_texts without pair_text"); return nullptr; }
  for (uint32_t i = 0; i < b->n_patterns; i++)
    if (b->pattern_offsets[i + 1] <= b->pattern_offsets[i]) { fail("nvbx_stream_create: empty pattern or decreasing offsets"); return nullptr; }
  for (uint32_t i = 0; i < b->n_texts; i++)
    if (b->text_offsets[i + 1] <= b->text_offsets[i]) { fail("nvbx_stream_create: empty text or decreasing offsets"); return nullptr; }
  nvbx_stream_t* s = new nvbx_stream();
  s->st = (cudaStream_t)cuda_stream;
  s->n_patterns = b->n_patterns; s->n_texts = b->n_texts; s->n_pairs = b->n_pairs; s->text_shared = text_shared;
  for (uint32_t k = 0; k < b->n_pairs; k++) {
    const uint32_t pi = b->pair_pattern ? b->pair_pattern[k] : k, ti = text_shared ? 0u : (b->pair_text ? b->pair_text[k] : k);
    if (pi >= b->n_patterns || ti >= b->n_texts) { fail("nvbx_stream_create: pair index out of range"); delete s; return nullptr; }
    const uint32_t m = b->pattern_offsets[pi + 1] - b->pattern_offsets[pi], n = b->text_offsets[ti + 1] - b->text_offsets[ti];
    s->max_m = std::max(s->max_m, m); s->max_n = std::max(s->max_n, n);
    s->cells += (uint64_t)m * n;
  }
  const size_t pb = b->pattern_offsets[b->n_patterns], tb = b->text_offsets[b->n_texts];
  const size_t ppw = (pb >> 4) + b->n_patterns + 2, tpw = (tb >> 4) + b->n_texts + 2;
  NVB_CHECKP(cudaMalloc(&s->pbytes, pb)); NVB_CHECKP(cudaMalloc(&s->tbytes, tb));
  NVB_CHECKP(cudaMalloc(&s->poffs, 4 * (size_t)(b->n_patterns + 1))); NVB_CHECKP(cudaMalloc(&s->toffs, 4 * (size_t)(b->n_texts + 1)));
  NVB_CHECKP(cudaMalloc(&s->ppk, 4 * ppw)); NVB_CHECKP(cudaMalloc(&s->tpk, 4 * tpw));
  NVB_CHECKP(cudaMalloc(&s->pflag, b->n_patterns)); NVB_CHECKP(cudaMalloc(&s->tflag, b->n_texts));
  NVB_CHECKP(cudaMalloc(&s->score, 4 * (size_t)b->n_pairs)); NVB_CHECKP(cudaMalloc(&s->sink_i, 4 * (size_t)b->n_pairs)); NVB_CHECKP(cudaMalloc(&s->sink_j, 4 * (size_t)b->n_pairs));
  NVB_CHECKP(cudaMalloc(&s->worklist, 4 * (size_t)b->n_pairs + 8)); NVB_CHECKP(cudaMalloc(&s->work_count, 4));
  NVB_CHECKP(cudaEventCreate(&s->e0)); NVB_CHECKP(cudaEventCreate(&s->e1));
  NVB_CHECKP(cudaMemcpyAsync(s->pbytes, b->patterns, pb, cudaMemcpyHostToDevice, s->st));
  NVB_CHECKP(cudaMemcpyAsync(s->tbytes, b->texts, tb, cudaMemcpyHostToDevice, s->st));
  NVB_CHECKP(cudaMemcpyAsync(s->poffs, b->pattern_offsets, 4 * (size_t)(b->n_patterns + 1), cudaMemcpyHostToDevice, s->st));
  NVB_CHECKP(cudaMemcpyAsync(s->toffs, b->text_offsets, 4 * (size_t)(b->n_texts + 1), cudaMemcpyHostToDevice, s->st));
  if (b->pair_pattern) { NVB_CHECKP(cudaMalloc(&s->pair_p, 4 * (size_t)b->n_pairs)); NVB_CHECKP(cudaMemcpyAsync(s->pair_p, b->pair_pattern, 4 * (size_t)b->n_pairs, cudaMemcpyHostToDevice, s->st)); }
  if (b->pair_text) { NVB_CHECKP(cudaMalloc(&s->pair_t, 4 * (size_t)b->n_pairs)); NVB_CHECKP(cudaMemcpyAsync(s->pair_t, b->pair_text, 4 * (size_t)b->n_pairs, cudaMemcpyHostToDevice, s->st)); }
  k_nvb_pack<<<(b->n_patterns + 7) / 8, 256, 0, s->st>>>(s->pbytes, s->poffs, b->n_patterns, b->alphabet, s->ppk, s->pflag);
  k_nvb_pack<<<(b->n_texts + 7) / 8, 256, 0, s->st>>>(s->tbytes, s->toffs, b->n_texts, b->alphabet, s->tpk, s->tflag);
  NVB_CHECKP(cudaGetLastError());
  g_launches += 2;
  return s;
}

int nvbx_stream_set_min_scores(nvbx_stream_t* s, const int32_t* per_job, int32_t uniform) {
  if (!s) return fail("nvbx_stream_set_min_scores: NULL stream");
  if (per_job) {
    if (!s->min_scores) NVB_CHECK(cudaMalloc(&s->min_scores, 4 * (size_t)s->n_pairs));
    NVB_CHECK(cudaMemcpyAsync(s->min_scores, per_job, 4 * (size_t)s->n_pairs, cudaMemcpyHostToDevice, s->st));
    NVB_CHECK(cudaStreamSynchronize(s->st));                  // the caller's array may be pageable and short-lived
    s->use_min_score = true;
  } else {
    if (s->min_scores) { cudaFree(s->min_scores); s->min_scores = nullptr; }
    s->use_min_score = uniform != INT32_MIN;
  }
  s->min_score = uniform;
  return 0;
}

int nvbx_stream_enact(nvbx_stream_t* s, const nvbx_aligner_t* al, int want_sinks) {
  if (!s || !al) return fail("nvbx_stream_enact: NULL argument");
  if (al->type < NVBX_GLOBAL || al->type > NVBX_SEMI_GLOBAL) return fail("nvbx_stream_enact: unknown alignment type");
  NvbArgs a{};
  a.pbytes = s->pbytes; a.poffs = s->poffs; a.ppk = s->ppk; a.pflag = s->pflag;
  a.tbytes = s->tbytes; a.toffs = s->toffs; a.tpk = s->tpk; a.tflag = s->tflag;
  a.pair_p = s->pair_p; a.pair_t = s->pair_t; a.text_shared = s->text_shared;
  switch (al->algo) {
    case NVBX_GOTOH: a.match = al->match; a.mismatch = al->mismatch; a.go_t = a.go_p = al->gap_open; a.ge_t = a.ge_p = al->gap_ext; a.band_shift = 3; break;
    case NVBX_SMITH_WATERMAN: a.match = al->match; a.mismatch = al->mismatch; a.go_t = a.ge_t = al->gap_open; a.go_p = a.ge_p = al->gap_ext; a.band_shift = 4; break;
    case NVBX_EDIT_DISTANCE: a.match = 0; a.mismatch = -1; a.go_t = a.ge_t = a.go_p = a.ge_p = -1; a.band_shift = 4; break;
    default: return fail("nvbx_stream_enact: unknown aligner");
  }
  a.score = s->score; a.sink_i = s->sink_i; a.sink_j = s->sink_j; a.worklist = s->worklist; a.work_count = s->work_count;

  // host-side range proof for the s16x2 path: every H + Go, E, F, X and diagonal sum of any job of the batch (pad rows / columns
  // included) stays inside [-31000, 32000]; substitution scores fit the byte tables (and stay above the -128 of GLOBAL's
  // virtual columns); symmetric gap costs with Go <= Ge <= 0 (the identity that takes H off the E chain)
  const long long mn = std::min(s->max_m, s->max_n), sum = (long long)s->max_m + s->max_n;
  const long long pos = std::max(0, std::max(a.match, a.mismatch));
  const long long hi = pos * mn + 2LL * std::abs(a.go_t) + 16;
  const long long lo = (al->type == NVBX_LOCAL) ? 3LL * a.go_t + std::min(a.mismatch, 0) + a.ge_t
                                                : 4LL * a.go_t + (long long)a.ge_t * (sum + 34) + std::min(a.mismatch, 0) + std::min(a.match, 0) - 128;
  const int tsh = chain_for(al->type) ? a.go_t : 0;                                        // the chain formulation's tables hold s - Go
  const bool fits = hi <= 32000 && lo >= -31000 && a.go_t == a.go_p && a.ge_t == a.ge_p && a.go_t <= a.ge_t && a.ge_t <= 0 && a.ge_t > -128 &&
                    a.go_t > -128 && a.mismatch <= 0 && a.match - tsh <= 127 && a.match - tsh > -128 && a.mismatch - tsh > -128;
  // the banded-16 int32 kernel needs ordered, non-positive costs and head room for its -2^20 front-column sentinel
  const bool wide = a.go_t <= a.ge_t && a.ge_t <= 0 && a.go_p <= a.ge_p && a.ge_p <= 0 && a.mismatch <= 0 &&
                    (sum + 34) * (long long)std::max(std::abs(a.ge_t), std::abs(a.ge_p)) + 4LL * std::max(std::abs(a.go_t), std::abs(a.go_p)) +
                        (long long)std::max(std::abs(a.match), std::abs(a.mismatch)) * (mn + 1) < (1LL << 28) &&
                    std::max(std::abs(a.ge_t), std::abs(a.ge_p)) < (1 << 19);
  // LOCAL sinks on the packed path: keys H * 16 + column must fit 16 bits, and strictly negative penalties keep pad cells from tying
  const bool local_keys = pos * mn <= 2047 && a.go_t < 0 && a.ge_t < 0 && a.mismatch < 0 && s->max_n < (1u << 20);
  // quad-strip kernel (nvb_strip.cuh): K2's blocking, no band-border scratch; scores of all three types, sinks of GLOBAL / SEMI_GLOBAL
  const bool strip = g_policy == 0 && wide && a.go_t == a.go_p && a.ge_t == a.ge_p &&
                     gb200::nvb_strip_eligible(al->type, a.match, -a.mismatch, -a.go_t, -a.ge_t, s->max_m, s->max_n, want_sinks);
  if (s->use_min_score) {
    // thresholds set: the band-of-8 int32 kernel, whose bands are the reference's (one thread per job, any alphabet, any costs)
    a.min_scores = s->min_scores; a.min_score = s->min_score;
    size_t chunk = ((size_t)1 << 28) / std::max<size_t>(1, s->max_n);
    chunk = std::max<size_t>(256, chunk & ~(size_t)255);
    chunk = std::min<size_t>(chunk, ((size_t)s->n_pairs + 255) & ~(size_t)255);
    const size_t need = chunk * s->max_n;
    if (need > s->scratch_elems) {
      if (s->scratch) cudaFree(s->scratch);
      s->scratch = nullptr; s->scratch_elems = 0;
      if (cudaMalloc(&s->scratch, need * sizeof(uint2)) != cudaSuccess) { cudaGetLastError(); return fail("nvbx_stream_enact: cannot allocate the band-border scratch (text too long for this batch)"); }
      s->scratch_elems = need;
    }
    a.scratch = s->scratch; a.stride = (uint32_t)chunk; a.rows_cap = s->max_n; a.sixteen = 16;
    NVB_CHECK(cudaEventRecord(s->e0, s->st));
    for (size_t first = 0; first < s->n_pairs; first += chunk) {
      a.first = (uint32_t)first; a.count = (uint32_t)std::min<size_t>(chunk, s->n_pairs - first);
      const uint32_t g = (a.count + kThreads - 1) / kThreads;
      const int type = al->type;
      if (want_sinks) {
        if (type == NVBX_GLOBAL) k_nvb_i32<NVBX_GLOBAL, true, true><<<g, kThreads, 0, s->st>>>(a, 0);
        else if (type == NVBX_LOCAL) k_nvb_i32<NVBX_LOCAL, true, true><<<g, kThreads, 0, s->st>>>(a, 0);
        else k_nvb_i32<NVBX_SEMI_GLOBAL, true, true><<<g, kThreads, 0, s->st>>>(a, 0);
      } else {
        if (type == NVBX_GLOBAL) k_nvb_i32<NVBX_GLOBAL, false, true><<<g, kThreads, 0, s->st>>>(a, 0);
        else if (type == NVBX_LOCAL) k_nvb_i32<NVBX_LOCAL, false, true><<<g, kThreads, 0, s->st>>>(a, 0);
        else k_nvb_i32<NVBX_SEMI_GLOBAL, false, true><<<g, kThreads, 0, s->st>>>(a, 0);
      }
      g_launches += 1;
      NVB_CHECK(cudaGetLastError());
    }
    NVB_CHECK(cudaEventRecord(s->e1, s->st));
    s->have_sinks = want_sinks != 0; s->enacted = true; s->have_tb = false;
    s->kernel = "int32-minscore";
    return 0;
  }
  const bool fast = strip || (fits && (!want_sinks || al->type != NVBX_LOCAL || local_keys) && (g_policy == 0 || g_policy == 2));
  a.sixteen = 16;
  gb200::NvbStripArgs sa{};
  if (strip) {
    sa.poffs = s->poffs; sa.ppk = s->ppk; sa.pflag = s->pflag; sa.toffs = s->toffs; sa.tpk = s->tpk; sa.tflag = s->tflag;
    sa.pair_p = s->pair_p; sa.pair_t = s->pair_t; sa.text_shared = s->text_shared;
    sa.match = a.match; sa.mismatch = -a.mismatch; sa.oe = -a.go_t; sa.ext = -a.ge_t;
    sa.score = s->score; sa.sink_i = s->sink_i; sa.sink_j = s->sink_j; sa.worklist = s->worklist; sa.work_count = s->work_count;
    sa.m8max = (s->max_m + 7u) & ~7u; sa.eight = 8; sa.one = 1;
  }

  // chunk the jobs so that the band-border scratch (8 B per text row per thread) stays below ~2 GB.  The fast path prefers the
  // fixed row stride (2 x kFixedStride jobs per launch) when the texts are short enough for it.
  const size_t cap_elems = (size_t)1 << 28;
  const bool fixed = fast && (size_t)s->max_n * kFixedStride <= cap_elems;
  size_t chunk = cap_elems / std::max<size_t>(1, s->max_n);
  chunk = std::max<size_t>(256, chunk & ~(size_t)255);
  chunk = std::min<size_t>(chunk, ((size_t)s->n_pairs + 255) & ~(size_t)255);
  if (fixed) chunk = 2 * (size_t)kFixedStride;
  const size_t need = (fixed ? (size_t)kFixedStride : chunk) * s->max_n;
  if (need > s->scratch_elems) {
    if (s->scratch) cudaFree(s->scratch);
    s->scratch = nullptr; s->scratch_elems = 0;
    if (cudaMalloc(&s->scratch, need * sizeof(uint2)) != cudaSuccess) { cudaGetLastError(); return fail("nvbx_stream_enact: cannot allocate the band-border scratch (text too long for this batch)"); }
    s->scratch_elems = need;
  }
  a.scratch = s->scratch; a.stride = (uint32_t)(fixed ? (size_t)kFixedStride : chunk); a.rows_cap = s->max_n;

  NVB_CHECK(cudaEventRecord(s->e0, s->st));
  bool used_i32 = !fast;
  for (size_t first = 0; first < s->n_pairs; first += chunk) {
    a.first = (uint32_t)first; a.count = (uint32_t)std::min<size_t>(chunk, s->n_pairs - first);
    const int type = al->type;
    if (fast) {
      NVB_CHECK(cudaMemsetAsync(s->work_count, 0, 4, s->st));
      // the worklist pass runs one job per THREAD on the scratch that was sized for the packed kernel's threads (two jobs each):
      // its grid is bounded by a.stride threads and loops (k_nvb_i32w), instead of indexing past the allocation
      const uint32_t items = (a.count + 1) / 2, g = (items + kThreads - 1) / kThreads, g2 = std::min<uint32_t>((a.count + kThreads - 1) / kThreads, a.stride / kThreads);
#define NVB_LAUNCH_FAST(T, S) do { \
        if (strip) { sa.first = a.first; sa.count = a.count; gb200::nvb_strip_launch(sa, T, S ? 1 : 0, s->st); } \
        else if (fixed) k_nvb_s16<T, true, S><<<g, kThreads, 0, s->st>>>(a); else k_nvb_s16<T, false, S><<<g, kThreads, 0, s->st>>>(a); \
        k_nvb_i32w<T, S><<<g2, kThreads, 0, s->st>>>(a, 1); } while (0)
      if (type == NVBX_LOCAL) { if (want_sinks) NVB_LAUNCH_FAST(NVBX_LOCAL, true); else NVB_LAUNCH_FAST(NVBX_LOCAL, false); }
      else if (type == NVBX_GLOBAL) { if (want_sinks) NVB_LAUNCH_FAST(NVBX_GLOBAL, true); else NVB_LAUNCH_FAST(NVBX_GLOBAL, false); }
      else { if (want_sinks) NVB_LAUNCH_FAST(NVBX_SEMI_GLOBAL, true); else NVB_LAUNCH_FAST(NVBX_SEMI_GLOBAL, false); }
#undef NVB_LAUNCH_FAST
      g_launches += 2;
    } else {
      const uint32_t g = (a.count + kThreads - 1) / kThreads;
#define NVB_LAUNCH_I32(T, S) do { if (wide) k_nvb_i32w<T, S><<<g, kThreads, 0, s->st>>>(a, 0); else k_nvb_i32<T, S><<<g, kThreads, 0, s->st>>>(a, 0); } while (0)
      if (want_sinks) {
        if (type == NVBX_GLOBAL) NVB_LAUNCH_I32(NVBX_GLOBAL, true); else if (type == NVBX_LOCAL) NVB_LAUNCH_I32(NVBX_LOCAL, true); else NVB_LAUNCH_I32(NVBX_SEMI_GLOBAL, true);
      } else {
        if (type == NVBX_GLOBAL) NVB_LAUNCH_I32(NVBX_GLOBAL, false); else if (type == NVBX_LOCAL) NVB_LAUNCH_I32(NVBX_LOCAL, false); else NVB_LAUNCH_I32(NVBX_SEMI_GLOBAL, false);
      }
#undef NVB_LAUNCH_I32
      g_launches += 1;
    }
    NVB_CHECK(cudaGetLastError());
  }
  NVB_CHECK(cudaEventRecord(s->e1, s->st));
  s->have_sinks = want_sinks != 0; s->enacted = true; s->have_tb = false;
  s->kernel = used_i32 ? (wide ? "int32" : "int32-any") : (strip ? "s16x2-strip" : "s16x2");
  return 0;
}

}  // extern "C"

// ---- banded scoring and traceback batches (kernels: nvb_tb.cuh) ------------------------------------------------------------------
static int base_args(nvbx_stream_t* s, const nvbx_aligner_t* al, NvbArgs& a, const char* who) {
  if (!s || !al) return fail(std::string(who) + ": NULL argument");
  if (al->type < NVBX_GLOBAL || al->type > NVBX_SEMI_GLOBAL) return fail(std::string(who) + ": unknown alignment type");
  a.pbytes = s->pbytes; a.poffs = s->poffs; a.ppk = s->ppk; a.pflag = s->pflag;
  a.tbytes = s->tbytes; a.toffs = s->toffs; a.tpk = s->tpk; a.tflag = s->tflag;
  a.pair_p = s->pair_p; a.pair_t = s->pair_t; a.text_shared = s->text_shared;
  switch (al->algo) {
    case NVBX_GOTOH: a.match = al->match; a.mismatch = al->mismatch; a.go_t = a.go_p = al->gap_open; a.ge_t = a.ge_p = al->gap_ext; a.band_shift = 3; break;
    case NVBX_SMITH_WATERMAN: a.match = al->match; a.mismatch = al->mismatch; a.go_t = a.ge_t = al->gap_open; a.go_p = a.ge_p = al->gap_ext; a.band_shift = 4; break;
    case NVBX_EDIT_DISTANCE: a.match = 0; a.mismatch = -1; a.go_t = a.ge_t = a.go_p = a.ge_p = -1; a.band_shift = 4; break;
    default: return fail(std::string(who) + ": unknown aligner");
  }
  a.score = s->score; a.sink_i = s->sink_i; a.sink_j = s->sink_j; a.worklist = s->worklist; a.work_count = s->work_count; a.sixteen = 16;
  return 0;
}
// banded costs: the pattern-consuming move (F) / the text-consuming move (E) / the row-zero init.  Gotoh uses one (open, extend)
// pair for all three; the banded Smith-Waterman prices the pattern-consuming move with deletion() and the text-consuming one with
// insertion() (sw_banded_inl.h:378-379,432-434 -- the opposite of its full-matrix code) and fills row zero with j * deletion().
static void banded_costs(const nvbx_aligner_t* al, const NvbArgs& a, TbArgs& tb) {
  if (al->algo == NVBX_GOTOH) { tb.f_o = tb.e_o = tb.init_o = a.go_t; tb.f_e = tb.e_e = tb.init_e = a.ge_t; tb.no_sink = 0; }
  else { tb.f_o = tb.f_e = tb.init_o = tb.init_e = a.go_t; tb.e_o = tb.e_e = a.go_p; tb.no_sink = 1; }
}
static int ensure_tb_outputs(nvbx_stream_t* s) {
  if (s->ops) return 0;                                               // all four exist, or none
  s->ops_stride = ((s->max_m + s->max_n + 64u) + 3u) & ~3u;
  const bool ok = cudaMalloc(&s->src_i, 4 * (size_t)s->n_pairs) == cudaSuccess && cudaMalloc(&s->src_j, 4 * (size_t)s->n_pairs) == cudaSuccess &&
                  cudaMalloc(&s->n_ops, 4 * (size_t)s->n_pairs) == cudaSuccess && cudaMalloc(&s->ops, (size_t)s->n_pairs * s->ops_stride) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    cudaFree(s->src_i); cudaFree(s->src_j); cudaFree(s->n_ops); cudaFree(s->ops);
    s->src_i = s->src_j = nullptr; s->n_ops = nullptr; s->ops = nullptr;
    return fail("nvbx traceback: cannot allocate the alignment / edit-op buffers");
  }
  return 0;
}
static int ensure_dirs(nvbx_stream_t* s, size_t words) {
  if (words <= s->dirs_words) return 0;
  if (s->dirs) cudaFree(s->dirs);
  s->dirs = nullptr; s->dirs_words = 0;
  if (cudaMalloc(&s->dirs, words * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); return fail("nvbx traceback: cannot allocate the flow matrix"); }
  s->dirs_words = words;
  return 0;
}
static bool band_ok(int b) { return b == 3 || b == 5 || b == 7 || b == 15 || b == 31; }
// jobs per traceback launch: as many as the flow-word budget holds, in warps of 32 slots, never more than the batch has (long
// strings: a handful of jobs must not reserve 128 slots of half a gigabyte each)
static size_t tb_chunk(size_t budget_words, size_t per_job, size_t n_pairs) {
  size_t chunk = std::max<size_t>(32, (budget_words / std::max<size_t>(1, per_job)) & ~(size_t)31);
  return std::min<size_t>(chunk, (n_pairs + 31) & ~(size_t)31);
}

template <int B, bool DIRS>
static void launch_banded(int type, uint32_t g, cudaStream_t st, const NvbArgs& a, const TbArgs& tb, int want_sinks) {
  if (type == NVBX_GLOBAL) k_nvb_banded<B, NVBX_GLOBAL, DIRS><<<g, kThreads, 0, st>>>(a, tb, want_sinks);
  else if (type == NVBX_LOCAL) k_nvb_banded<B, NVBX_LOCAL, DIRS><<<g, kThreads, 0, st>>>(a, tb, want_sinks);
  else k_nvb_banded<B, NVBX_SEMI_GLOBAL, DIRS><<<g, kThreads, 0, st>>>(a, tb, want_sinks);
}
template <bool DIRS>
static void launch_banded_b(int band, int type, uint32_t g, cudaStream_t st, const NvbArgs& a, const TbArgs& tb, int want_sinks) {
  switch (band) {
    case 3: launch_banded<3, DIRS>(type, g, st, a, tb, want_sinks); break;
    case 5: launch_banded<5, DIRS>(type, g, st, a, tb, want_sinks); break;
    case 7: launch_banded<7, DIRS>(type, g, st, a, tb, want_sinks); break;
    case 15: launch_banded<15, DIRS>(type, g, st, a, tb, want_sinks); break;
    default: launch_banded<31, DIRS>(type, g, st, a, tb, want_sinks); break;
  }
}

extern "C" {

int nvbx_stream_enact_banded(nvbx_stream_t* s, const nvbx_aligner_t* al, int band_len, int want_sinks) {
  NvbArgs a{}; TbArgs tb{};
  if (base_args(s, al, a, "nvbx_stream_enact_banded") != 0) return -1;
  if (!band_ok(band_len)) return fail("nvbx_stream_enact_banded: band_len must be 3, 5, 7, 15 or 31 (the reference's instantiations)");
  banded_costs(al, a, tb);
  NVB_CHECK(cudaEventRecord(s->e0, s->st));
  a.first = 0; a.count = s->n_pairs;
  launch_banded_b<false>(band_len, al->type, (a.count + kThreads - 1) / kThreads, s->st, a, tb, want_sinks ? 1 : 0);
  NVB_CHECK(cudaGetLastError());
  g_launches += 1;
  NVB_CHECK(cudaEventRecord(s->e1, s->st));
  s->have_sinks = want_sinks != 0; s->enacted = true; s->have_tb = false;
  s->kernel = "banded-int32";
  return 0;
}

int nvbx_stream_traceback(nvbx_stream_t* s, const nvbx_aligner_t* al, int band_len) {
  NvbArgs a{}; TbArgs tb{};
  if (base_args(s, al, a, "nvbx_stream_traceback") != 0) return -1;
  if (band_len != 0 && !band_ok(band_len)) return fail("nvbx_stream_traceback: band_len must be 0 (full matrix), 3, 5, 7, 15 or 31");
  if (ensure_tb_outputs(s) != 0) return -1;
  tb.src_i = s->src_i; tb.src_j = s->src_j; tb.n_ops = s->n_ops; tb.ops = s->ops; tb.ops_stride = s->ops_stride;
  size_t budget_words = (size_t)3 << 29;                                                   // <= 6 GB of flow words per chunk
  if (const char* e = getenv("NVBX_TB_BUDGET_WORDS")) budget_words = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 10));   // tests: force several chunks
  NVB_CHECK(cudaEventRecord(s->e0, s->st));
  if (band_len) {
    banded_costs(al, a, tb);
    const size_t per_job = (size_t)s->max_m * ((band_len + 7) / 8);
    size_t chunk = tb_chunk(budget_words, per_job, s->n_pairs);
    if (ensure_dirs(s, per_job * chunk) != 0) return -1;
    tb.dirs = s->dirs; tb.slots = chunk; tb.rows_cap = s->max_m;
    for (size_t first = 0; first < s->n_pairs; first += chunk) {
      a.first = (uint32_t)first; a.count = (uint32_t)std::min<size_t>(chunk, s->n_pairs - first);
      const uint32_t g = (a.count + kThreads - 1) / kThreads;
      launch_banded_b<true>(band_len, al->type, g, s->st, a, tb, 1);
      switch (band_len) {
        case 3: k_nvb_banded_walk<3><<<g, kThreads, 0, s->st>>>(a, tb, al->type); break;
        case 5: k_nvb_banded_walk<5><<<g, kThreads, 0, s->st>>>(a, tb, al->type); break;
        case 7: k_nvb_banded_walk<7><<<g, kThreads, 0, s->st>>>(a, tb, al->type); break;
        case 15: k_nvb_banded_walk<15><<<g, kThreads, 0, s->st>>>(a, tb, al->type); break;
        default: k_nvb_banded_walk<31><<<g, kThreads, 0, s->st>>>(a, tb, al->type); break;
      }
      NVB_CHECK(cudaGetLastError());
      g_launches += 2;
    }
    s->kernel = "banded-int32+tb";
  } else {
    const size_t nbands = (s->max_m + 7) / 8, per_job = nbands * s->max_n;
    size_t chunk = tb_chunk(budget_words, per_job, s->n_pairs);
    if (ensure_dirs(s, per_job * chunk) != 0) return -1;
    const size_t need = ((chunk + kThreads - 1) / kThreads) * kThreads * s->max_n;         // band-border scratch, [CTA][text row][thread]
    if (need > s->scratch_elems) {
      if (s->scratch) cudaFree(s->scratch);
      s->scratch = nullptr; s->scratch_elems = 0;
      if (cudaMalloc(&s->scratch, need * sizeof(uint2)) != cudaSuccess) { cudaGetLastError(); return fail("nvbx_stream_traceback: cannot allocate the band-border scratch"); }
      s->scratch_elems = need;
    }
    a.scratch = s->scratch; a.stride = (uint32_t)chunk; a.rows_cap = s->max_n;
    tb.dirs = s->dirs; tb.slots = chunk; tb.rows_cap = s->max_n;
    for (size_t first = 0; first < s->n_pairs; first += chunk) {
      a.first = (uint32_t)first; a.count = (uint32_t)std::min<size_t>(chunk, s->n_pairs - first);
      const uint32_t g = (a.count + kThreads - 1) / kThreads;
      if (al->type == NVBX_GLOBAL) k_nvb_tb_fwd<NVBX_GLOBAL><<<g, kThreads, 0, s->st>>>(a, tb);
      else if (al->type == NVBX_LOCAL) k_nvb_tb_fwd<NVBX_LOCAL><<<g, kThreads, 0, s->st>>>(a, tb);
      else k_nvb_tb_fwd<NVBX_SEMI_GLOBAL><<<g, kThreads, 0, s->st>>>(a, tb);
      k_nvb_tb_walk<<<g, kThreads, 0, s->st>>>(a, tb, al->type);
      NVB_CHECK(cudaGetLastError());
      g_launches += 2;
    }
    s->kernel = "int32+tb";
  }
  NVB_CHECK(cudaEventRecord(s->e1, s->st));
  s->have_sinks = true; s->enacted = true; s->have_tb = true;
  return 0;
}

uint32_t nvbx_stream_ops_stride(const nvbx_stream_t* s) { return s ? (s->ops_stride ? s->ops_stride : (((s->max_m + s->max_n + 64u) + 3u) & ~3u)) : 0; }

int nvbx_stream_alignments(nvbx_stream_t* s, nvbx_alignment_t* alns, uint8_t* ops) {
  if (!s || !alns) return fail("nvbx_stream_alignments: NULL argument");
  if (!s->have_tb) return fail("nvbx_stream_alignments: no traceback was enacted");
  const size_t n = s->n_pairs;
  std::vector<int32_t> h(5 * n); std::vector<uint32_t> hn(n);
  NVB_CHECK(cudaMemcpyAsync(h.data(), s->score, 4 * n, cudaMemcpyDeviceToHost, s->st));
  NVB_CHECK(cudaMemcpyAsync(h.data() + n, s->src_i, 4 * n, cudaMemcpyDeviceToHost, s->st));
  NVB_CHECK(cudaMemcpyAsync(h.data() + 2 * n, s->src_j, 4 * n, cudaMemcpyDeviceToHost, s->st));
  NVB_CHECK(cudaMemcpyAsync(h.data() + 3 * n, s->sink_i, 4 * n, cudaMemcpyDeviceToHost, s->st));
  NVB_CHECK(cudaMemcpyAsync(h.data() + 4 * n, s->sink_j, 4 * n, cudaMemcpyDeviceToHost, s->st));
  NVB_CHECK(cudaMemcpyAsync(hn.data(), s->n_ops, 4 * n, cudaMemcpyDeviceToHost, s->st));
  if (ops) NVB_CHECK(cudaMemcpyAsync(ops, s->ops, n * s->ops_stride, cudaMemcpyDeviceToHost, s->st));
  NVB_CHECK(cudaStreamSynchronize(s->st));
  std::vector<uint32_t> hm(s->n_patterns + 1);
  NVB_CHECK(cudaMemcpy(hm.data(), s->poffs, 4 * (size_t)(s->n_patterns + 1), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> pp;
  if (s->pair_p) { pp.resize(n); NVB_CHECK(cudaMemcpy(pp.data(), s->pair_p, 4 * n, cudaMemcpyDeviceToHost)); }
  for (size_t k = 0; k < n; k++) {
    nvbx_alignment_t& o = alns[k];
    o.score = h[k]; o.source_text = h[n + k]; o.source_pattern = h[2 * n + k]; o.sink_text = h[3 * n + k]; o.sink_pattern = h[4 * n + k];
    o.n_ops = hn[k];
    const uint32_t pi = s->pair_p ? pp[k] : (uint32_t)k, m = hm[pi + 1] - hm[pi];
    const bool none = o.sink_text < 0 || o.sink_pattern < 0;
    o.clip_end = none ? 0u : m - (uint32_t)o.sink_pattern;                                  // backtracer.clip( pattern.length() - sink.y )
    o.clip_begin = none ? 0u : (uint32_t)o.source_pattern;                                  // backtracer.clip( source.y )
  }
  return 0;
}

int nvbx_stream_results(nvbx_stream_t* s, int32_t* scores, int32_t* sink_text, int32_t* sink_pattern) {
  if (!s || !scores) return fail("nvbx_stream_results: NULL argument");
  if (!s->enacted) return fail("nvbx_stream_results: nothing was enacted");
  if ((sink_text || sink_pattern) && !s->have_sinks) return fail("nvbx_stream_results: sinks were not computed (want_sinks = 0)");
  NVB_CHECK(cudaMemcpyAsync(scores, s->score, 4 * (size_t)s->n_pairs, cudaMemcpyDeviceToHost, s->st));
  if (sink_text) NVB_CHECK(cudaMemcpyAsync(sink_text, s->sink_i, 4 * (size_t)s->n_pairs, cudaMemcpyDeviceToHost, s->st));
  if (sink_pattern) NVB_CHECK(cudaMemcpyAsync(sink_pattern, s->sink_j, 4 * (size_t)s->n_pairs, cudaMemcpyDeviceToHost, s->st));
  NVB_CHECK(cudaStreamSynchronize(s->st));
  return 0;
}

const int32_t* nvbx_stream_scores_device(nvbx_stream_t* s) { return s ? s->score : nullptr; }
float nvbx_stream_last_ms(nvbx_stream_t* s) {
  if (!s || !s->enacted) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(s->e1) != cudaSuccess || cudaEventElapsedTime(&ms, s->e0, s->e1) != cudaSuccess) return -1.f;
  return ms;
}
uint64_t nvbx_stream_cells(const nvbx_stream_t* s) { return s ? s->cells : 0; }
const char* nvbx_stream_kernel(const nvbx_stream_t* s) { return s ? s->kernel.c_str() : "none"; }

int nvbx_batched_score_host(const nvbx_aligner_t* al, const nvbx_batch_t* b, int32_t* scores, int32_t* sink_text, int32_t* sink_pattern, float ms[3]) {
  cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
  for (auto& x : e) if (cudaEventCreate(&x) != cudaSuccess) return fail("nvbx_batched_score_host: no CUDA device");
  int rc = -1;
  cudaEventRecord(e[0], 0);
  nvbx_stream_t* s = nvbx_stream_create(b, nullptr);
  if (s) {
    cudaEventRecord(e[1], 0);
    rc = nvbx_stream_enact(s, al, (sink_text || sink_pattern) ? 1 : 0);
    if (rc == 0) {
      cudaEventRecord(e[2], 0);
      rc = nvbx_stream_results(s, scores, sink_text, sink_pattern);
      cudaEventRecord(e[3], 0);
      cudaEventSynchronize(e[3]);
      if (rc == 0 && ms) for (int k = 0; k < 3; k++) cudaEventElapsedTime(&ms[k], e[k], e[k + 1]);
    }
    nvbx_stream_destroy(s);
  }
  for (auto& x : e) cudaEventDestroy(x);
  return rc;
}

}  // extern "C"
