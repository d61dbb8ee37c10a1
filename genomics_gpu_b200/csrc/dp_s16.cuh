// dp_s16.cuh -- the s16x2 DPX fast path (K2): inter-task kernels for short-read shapes.
//
// One thread = one pair.  The two 16-bit lanes of every register hold TWO CONSECUTIVE 8-ROW STRIPS
// OF THE SAME PAIR, skewed by one query column: at step k lane L (bits 15:0) computes column k of
// strip 2j while lane H (bits 31:16) computes column k-1 of strip 2j+1, so lane L's bottom row feeds
// lane H's top row through registers one step later and only one (H,E) word per column per 16 target
// rows goes through shared memory.  Per cell pair the work is split over BOTH integer pipes of the SM
// (measured on B200, profiles/dpx_microbench_r01.json: VIADD.16x2 / IMAD issue on the FMA pipe
// concurrently with the DPX/ALU pipe):
//     ALU pipe : PRMT (substitution lookup), VIMNMX3.S16x2.RELU (H), 2 x VIADDMNMX.S16x2 (E, F)
//     FMA pipe : 2 x VIADD.16x2 (D = diag + s, D - oe), IMAD (argmax key)
// Substitution scores come from a per-row 4-byte table (one byte per query base A/C/G/T, sign
// replicated to 16 bits by PRMT) selected by a per-column selector byte that is precomputed once per
// pair into shared memory.  The argmax with the reference's (strip, column, row) tie order is tracked
// without branches: key = H*8 + (7-row) per cell, a VIMNMX3 tree per column, and a predicated record of
// the column -- see DESIGN.md "K2".
//
// Exactness: identical results to the general kernels / the reference for every pair the fast path
// accepts.  Pairs it cannot represent (a base that is not A/C/G/T inside the true length, e.g. N) are
// pushed to a worklist and recomputed by the general int32 kernel in the same stream.
#pragma once
#include "dp_common.cuh"
#include "dp_generic.cuh"
#include "dp_launch.h"
#include <type_traits>
#include <stdlib.h>
#pragma nv_diag_suppress 128   // "loop is not reachable": k_local_s16<START_TB> returns before the score-only sweeps

namespace gb200 {

#ifndef GB200_FAST_THREADS
#define GB200_FAST_THREADS 128
#endif
constexpr int kFastThreads = GB200_FAST_THREADS;
// launch bounds of the s16x2 kernels; GB200_FAST_MINBLOCKS (experiments): promise the compiler that many resident CTAs per SM.
// Measured on C2 (262144 pairs; ms for local / local+start / semi / semi+start / global): default 4.28 / 4.95 / 3.72 / 4.54 / 3.62,
// 2 CTAs (more registers) 4.29 / 4.92 / 3.99 / 4.81 / 3.66, 3 CTAs 4.28 / 4.95 / 3.82 / 4.62 / 3.66 -- the default stays.
#ifdef GB200_FAST_MINBLOCKS
#define GB200_FAST_LB __launch_bounds__(kFastThreads, GB200_FAST_MINBLOCKS)
#else
#define GB200_FAST_LB __launch_bounds__(kFastThreads)
#endif
constexpr uint32_t kSlotPad = 8;          // dummy column slots: 2 below column 0 (late lanes' phantom columns), 6 above the last column
constexpr uint32_t kFastFallbackSlots = 8192;

struct FastArgs {
  DpArgs a;
  uint32_t q8max;          // columns reserved per thread in shared memory (multiple of 8)
  uint32_t* worklist; uint32_t* work_count;     // pairs with a base outside A/C/G/T (filled by the main launch; counts: work_count[0], [1])
  uint32_t* worklist2;                          // pairs the N-aware pass cannot represent either (codes other than A/C/G/T/N): general kernel
  const uint32_t* flags;                        // k_classify: 0 = pure A/C/G/T pair (main launch), bit 0: N (worklist), bit 1: general kernel (worklist2);
                                                // nullptr: the host packer saw nothing but A/C/G/T in this batch, there is nothing to skip
  uint32_t eight;          // == 8, passed at run time so the key multiply stays an IMAD (FMA pipe)
  uint32_t thirtytwo;      // == 32, same for the grouped arg-max keys
  uint32_t one;            // == 1, run-time so that the sweep's trip count (columns + 1) is opaque to the unroller
  uint32_t mm, sixteen;    // == 0x00204081 / 16, run-time so the direction-nibble gather stays on IMAD (FMA pipe)
};

__device__ __forceinline__ uint32_t splat16(int v) { const uint32_t u = (uint32_t)v & 0xFFFFu; return u | (u << 16); }

// per-row substitution table: byte i = score of query base i against this row's target base
__device__ __forceinline__ uint32_t row_table(uint32_t idx, uint32_t mis4, uint32_t delta) {
  return (idx < 4) ? (mis4 ^ (delta << (8 * idx))) : mis4;
}

// N-aware variant (NMASK): a real N in the query scores n_score against everything (gasal_kernels.h:49-51).  The PRMT table has no
// room for a fifth column, so the N columns are marked in the selector array (0xFF) and one LOP3 per cell pair overrides the
// looked-up score with the N score under a per-lane column mask; N rows of the target are folded into the row tables.
__device__ __forceinline__ uint32_t mask_sel(uint32_t s2, uint32_t m, uint32_t nsc2) { return (s2 & ~m) | (nsc2 & m); }
__device__ __forceinline__ uint32_t ncol_mask16(uint32_t v) { return (v == 0xFFu) ? 0xFFFFu : 0u; }
constexpr uint32_t kSelN = 0xFFu;

struct LaneBest { uint32_t bestK7, keyL, keyH, posL, posH; };

// bestK7 = max.s16x2(bestK7, c7); lanes that strictly improved record (cm, pos).  One asm block so that
// ptxas fuses the per-lane compares into the predicate outputs of VIMNMX.S16x2 (the CUDA header's
// __vibmax_s16x2 lost its predicate uses in partially unrolled loops with nvcc 12.9).
#ifndef GB200_PRUNE_ARGMAX
#define GB200_PRUNE_ARGMAX 1
#endif
#ifndef GB200_PRUNE_REVERSE
#define GB200_PRUNE_REVERSE 1
#endif
#ifndef GB200_TBB_UNROLL
#define GB200_TBB_UNROLL 2          // step loop of the banded direction sweep: 2 = +24 % (shift registers and diagonal rotation become renaming); 4 = -35 % (instruction cache)
#endif
#ifndef GB200_TB_UNROLL
#define GB200_TB_UNROLL 2          // same for the unbanded direction sweep: +8 %
#endif
#define GB200_PRAGMA_(x) _Pragma(#x)
#define GB200_UNROLL_(n) GB200_PRAGMA_(unroll n)
#define GB200_TBB_UNROLL_PRAGMA GB200_UNROLL_(GB200_TBB_UNROLL)
#define GB200_TB_UNROLL_PRAGMA GB200_UNROLL_(GB200_TB_UNROLL)
#ifndef GB200_BSC_UNROLL
#define GB200_BSC_UNROLL 2          // step loop of the score-only banded sweep
#endif
#define GB200_BSC_UNROLL_PRAGMA GB200_UNROLL_(GB200_BSC_UNROLL)
#ifndef GB200_PRUNE_REVERSE_LOW
#define GB200_PRUNE_REVERSE_LOW 1
#endif
#ifndef GB200_PRUNE_RISKY_LANES
#define GB200_PRUNE_RISKY_LANES 4
#endif
#ifndef GB200_TRACK_IMAD
#define GB200_TRACK_IMAD 1
#endif
__device__ __forceinline__ void track_update(LaneBest& lb, uint32_t c7, uint32_t cm, uint32_t pos, uint32_t one = 1u) {
#if GB200_TRACK_IMAD
  // the four predicated records as multiply-adds by a run-time 1: ptxas keeps them as predicated IMADs on the FMA pipe instead of
  // SELs on the ALU pipe, which is the pipe that bounds these kernels
  asm volatile(
      "{\n\t"
      ".reg .pred pl, ph;\n\t"
      ".reg .s16 n0, n1, o0, o1;\n\t"
      ".reg .b32 nb;\n\t"
      "max.s16x2 nb, %0, %5;\n\t"
      "mov.b32 {n0, n1}, nb;\n\t"
      "mov.b32 {o0, o1}, %0;\n\t"
      "setp.eq.s16 pl, n0, o0;\n\t"
      "setp.eq.s16 ph, n1, o1;\n\t"
      "mov.b32 %0, nb;\n\t"
      "@!pl mad.lo.u32 %1, %6, %8, 0;\n\t"
      "@!pl mad.lo.u32 %3, %7, %8, 0;\n\t"
      "@!ph mad.lo.u32 %2, %6, %8, 0;\n\t"
      "@!ph mad.lo.u32 %4, %7, %8, 0;\n\t"
      "}"
      : "+r"(lb.bestK7), "+r"(lb.keyL), "+r"(lb.keyH), "+r"(lb.posL), "+r"(lb.posH)
      : "r"(c7), "r"(cm), "r"(pos), "r"(one));
  return;
#endif
  asm volatile(
      "{\n\t"
      ".reg .pred pl, ph;\n\t"
      ".reg .s16 n0, n1, o0, o1;\n\t"
      ".reg .b32 nb;\n\t"
      "max.s16x2 nb, %0, %5;\n\t"
      "mov.b32 {n0, n1}, nb;\n\t"
      "mov.b32 {o0, o1}, %0;\n\t"
      "setp.eq.s16 pl, n0, o0;\n\t"
      "setp.eq.s16 ph, n1, o1;\n\t"
      "mov.b32 %0, nb;\n\t"
      "@!pl mov.b32 %1, %6;\n\t"
      "@!pl mov.b32 %3, %7;\n\t"
      "@!ph mov.b32 %2, %6;\n\t"
      "@!ph mov.b32 %4, %7;\n\t"
      "}"
      : "+r"(lb.bestK7), "+r"(lb.keyL), "+r"(lb.keyH), "+r"(lb.posL), "+r"(lb.posH)
      : "r"(c7), "r"(cm), "r"(pos), "r"(one));
}

// One double-strip sweep of the D-based local recurrence (reference local_kernel_template.h:19-30).
//   HE  : shared (H | E<<16) per column, slot c+1 holds column c; slot 0 / slot ncol+1 are dummies
//   V   : shared selector byte per column, same slot numbering
//   rowA/rowB : substitution tables of the 8 rows of lane L's strip / lane H's strip
//   pos_base  : (double-strip index << 16), OR-ed with the step to form the recorded position
__device__ __forceinline__ void local_sweep(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, int nsteps,
                                            const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                            uint32_t neg_oe2, uint32_t neg_ext2, uint32_t eight, uint32_t pos_base, LaneBest& lb) {
  uint32_t p[8], f[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { p[r] = 0; f[r] = 0; }
  uint32_t h8 = 0, eo = 0;                   // bottom-row H and outgoing E of the previous step (lane L feeds lane H)
  uint32_t prevsh = 0xCC00u;                 // lane-H selector of the phantom column -1: score 0 / -1
  uint32_t he = HE[kFastThreads];            // slot 1 = column 0
  uint32_t v = V[kFastThreads];
  uint32_t pos = pos_base;
  for (int k = 0; k < nsteps; k++) {
    const uint32_t sel = v | prevsh;
    prevsh = v * 256u + 0x4400u;
    // lane L takes (H,E) of the strip above from shared memory, lane H takes lane L's bottom row of the previous step
    uint32_t up = prmt(he, h8, 0x5410);
    uint32_t e = prmt(he, eo, 0x5432);
    // prefetch the next step's inputs
    he = HE[(k + 2) * kFastThreads];
    v = V[(k + 2) * kFastThreads];
    uint32_t key[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint32_t s2 = prmt(rowA[r], rowB[r], sel);
      const uint32_t d = __vadd2(p[r], s2);
      const uint32_t dm = __vadd2(d, neg_oe2);
      const uint32_t h = __vimax3_s16x2_relu(d, f[r], e);
      f[r] = __viaddmax_s16x2(f[r], neg_ext2, dm);
      e = __viaddmax_s16x2(e, neg_ext2, dm);
      key[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
      p[r] = up; up = h;
    }
    // lane H finished column k-1 of its strip: park (H,E) for the next double-strip
    HE[k * kFastThreads] = prmt(up, e, 0x7632);
    h8 = up; eo = e;
    // branch-free argmax: column max of the keys, strict improvement test per lane on (score<<3 | 7)
    uint32_t cm = __vimax3_s16x2(key[0], key[1], key[2]);
    cm = __vimax3_s16x2(cm, key[3], key[4]);
    cm = __vimax3_s16x2(cm, key[5], key[6]);
    cm = __vmaxs2(cm, key[7]);
    const uint32_t c7 = cm | 0x00070007u;
    track_update(lb, c7, cm, pos);
    pos++;
  }
}

// decode the two lanes' records into the reference's (score, column, row) winner
struct Winner { int score, col, row; };   // row = absolute row index in sweep order (strip*8 + r)
__device__ __forceinline__ Winner pick_winner(const LaneBest& lb) {
  const int sL = (int)(short)(lb.bestK7 & 0xFFFFu) >> 3, sH = (int)(short)(lb.bestK7 >> 16) >> 3;
  Winner w; w.score = 0; w.col = 0; w.row = 0;
  const int stripL = 2 * (int)(lb.posL >> 16), stripH = 2 * (int)(lb.posH >> 16) + 1;
  const bool takeL = (sL > sH) || (sL == sH && stripL < stripH);
  if (takeL) {
    if (sL > 0) { w.score = sL; w.col = (int)(lb.posL & 0xFFFFu); w.row = stripL * 8 + (7 - (int)(lb.keyL & 7u)); }
  } else {
    if (sH > 0) { w.score = sH; w.col = (int)(lb.posH & 0xFFFFu) - 1; w.row = stripH * 8 + (7 - (int)((lb.keyH >> 16) & 7u)); }
  }
  return w;
}


// ---- quad-strip variant: FOUR consecutive strips per thread and sweep -------------------------------------
// Register set A holds strips 4j (lane L) and 4j+1 (lane H), set B strips 4j+2 / 4j+3; the four strips are skewed
// by one query column each (A.L col k, A.H col k-1, B.L col k-2, B.H col k-3), so B.L's top row is A.H's bottom row
// of the previous step.  Two independent dependency chains per thread (twice the ILP at the 8-9 resident warps per
// SM that shared memory allows), and the per-column shared-memory round trip is paid once per 32 target rows.
struct LaneBest4 { LaneBest a, b; };

__device__ __forceinline__ uint32_t key_tree(const uint32_t (&key)[8]) {
  uint32_t cm = __vimax3_s16x2(key[0], key[1], key[2]);
  cm = __vimax3_s16x2(cm, key[3], key[4]);
  cm = __vimax3_s16x2(cm, key[5], key[6]);
  return __vmaxs2(cm, key[7]);
}

template <bool NMASK = false>
__device__ __forceinline__ void local_sweep4(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, int nsteps,
                                             const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                             const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                             uint32_t neg_oe2, uint32_t neg_ext2, uint32_t eight, uint32_t one, uint32_t pos_base, LaneBest4& lb,
                                             uint32_t nsc2 = 0) {
  uint32_t pA[8], fA[8], pB[8], fB[8], keyA[8], keyB[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { pA[r] = 0; fA[r] = 0; pB[r] = 0; fB[r] = 0; keyA[r] = 0; keyB[r] = 0; }
  uint32_t h8A = 0, eoA = 0, h8B = 0, eoB = 0;
  uint32_t prevsh = 0xCC00u;
  uint32_t selB = 0xCCCCu, selA_prev = 0xCCCCu;   // phantom columns: score 0 / -1
  uint32_t he = HE[kFastThreads];                  // slot 1 = column 0
  uint32_t v = V[kFastThreads];
  uint32_t pos = pos_base;                         // position code of the step whose keys are in keyA / keyB
  uint32_t ncm_prev = 0, mB = 0, mA_prev = 0;      // NMASK: N-column masks, rotated like the selectors
  for (int k = 0; k < nsteps; k++) {
    const uint32_t selA = v | prevsh;
    prevsh = v * 256u + 0x4400u;
    uint32_t mA = 0;
    if (NMASK) { const uint32_t ncm = ncol_mask16(v); mA = ncm | (ncm_prev << 16); ncm_prev = ncm; }
    uint32_t upA = prmt(he, h8A, 0x5410);
    uint32_t eA = prmt(he, eoA, 0x5432);
    uint32_t upB = prmt(h8A, h8B, 0x5432);
    uint32_t eB = prmt(eoA, eoB, 0x5432);
    he = HE[(k + 2) * kFastThreads];
    v = V[(k + 2) * kFastThreads];
    // branch-free argmax of the PREVIOUS step's keys (software-pipelined: this dependent chain of 7 ALU ops has no
    // input from the current step, so it overlaps the row updates below instead of trailing them)
    {
      const uint32_t cm = key_tree(keyA);
      track_update(lb.a, cm | 0x00070007u, cm, pos, one);
      const uint32_t cn = key_tree(keyB);
      track_update(lb.b, cn | 0x00070007u, cn, pos, one);
      pos = pos_base + (uint32_t)k;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
      {
        uint32_t s2 = prmt(rowA[r], rowB[r], selA);
        if (NMASK) s2 = mask_sel(s2, mA, nsc2);
        const uint32_t d = __vadd2(pA[r], s2);
        const uint32_t dm = __vadd2(d, neg_oe2);
        const uint32_t h = __vimax3_s16x2_relu(d, fA[r], eA);
        fA[r] = __viaddmax_s16x2(fA[r], neg_ext2, dm);
        eA = __viaddmax_s16x2(eA, neg_ext2, dm);
        keyA[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
        pA[r] = upA; upA = h;
      }
      {
        uint32_t s2 = prmt(rowC[r], rowD[r], selB);
        if (NMASK) s2 = mask_sel(s2, mB, nsc2);
        const uint32_t d = __vadd2(pB[r], s2);
        const uint32_t dm = __vadd2(d, neg_oe2);
        const uint32_t h = __vimax3_s16x2_relu(d, fB[r], eB);
        fB[r] = __viaddmax_s16x2(fB[r], neg_ext2, dm);
        eB = __viaddmax_s16x2(eB, neg_ext2, dm);
        keyB[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
        pB[r] = upB; upB = h;
      }
    }
    // B.H finished column k-3 (slot k-2): park (H,E) for the next quad-strip
    HE[(k - 2) * kFastThreads] = prmt(upB, eB, 0x7632);
    h8A = upA; eoA = eA; h8B = upB; eoB = eB;
    selB = selA_prev; selA_prev = selA;
    if (NMASK) { mB = mA_prev; mA_prev = mA; }
  }
  {
    const uint32_t cm = key_tree(keyA);
    track_update(lb.a, cm | 0x00070007u, cm, pos, one);
    const uint32_t cn = key_tree(keyB);
    track_update(lb.b, cn | 0x00070007u, cn, pos, one);
  }
}

// ---- grouped arg-max (G4): the column maxima of FOUR consecutive steps are folded into one running maximum per lane before the
// strict-improvement test and the record, so the compare-and-record (LOP3 + VIMNMX + four selects per register set) runs once per
// group instead of once per step.  Keys carry the step's phase inside the group above the row: key = H * 32 + (3 - phase) * 8 +
// (7 - row), so the maximum still picks the earliest column, then the lowest row, of the best score; folding the running maximum
// into the per-step VIMNMX3 tree is free (nine inputs = four instructions, like eight).  The step loop is written as groups of four
// unrolled steps: no per-step branch.  Needs score * 32 + 31 < 2^15 (host check), else the per-step variant above runs.
__device__ __forceinline__ uint32_t key_tree9(uint32_t g, const uint32_t (&key)[8]) {
  uint32_t cm = __vimax3_s16x2(g, key[0], key[1]);
  cm = __vimax3_s16x2(cm, key[2], key[3]);
  cm = __vimax3_s16x2(cm, key[4], key[5]);
  return __vimax3_s16x2(cm, key[6], key[7]);
}

// Score-bound pruning of the arg-max (exact): a local alignment that ends in query column q has at most q + 1 matches, so every
// cell of that column has H <= match * (q + 1).  The forward pass only records STRICT improvements of the running best, and the
// start pass only needs the cells that reach the forward score -- so the columns with match * (q + 1) <= bound can never be
// recorded and their steps run without keys, trees and compare-and-record (`gfree` leading groups of four steps; the lanes of a
// step sit on columns <= the step index).  After a read has been found, almost the whole sweep runs in this tracking-free loop.
template <bool NMASK = false>
__device__ __forceinline__ void local_sweep4g(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, int ngroups, int gfree,
                                              const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                              const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                              uint32_t neg_oe2, uint32_t neg_ext2, uint32_t thirtytwo, uint32_t one, uint32_t pos_base, LaneBest4& lb,
                                              uint32_t nsc2 = 0) {
  uint32_t pA[8], fA[8], pB[8], fB[8], keyA[8], keyB[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { pA[r] = 0; fA[r] = 0; pB[r] = 0; fB[r] = 0; keyA[r] = 0; keyB[r] = 0; }
  uint32_t h8A = 0, eoA = 0, h8B = 0, eoB = 0;
  uint32_t prevsh = 0xCC00u;
  uint32_t selB = 0xCCCCu, selA_prev = 0xCCCCu;   // phantom columns: score 0 / -1
  uint32_t he = HE[kFastThreads];                  // slot 1 = column 0
  uint32_t v = V[kFastThreads];
  uint32_t gA = 0, gB = 0;                         // running maxima of the current group
  uint32_t ncm_prev = 0, mB = 0, mA_prev = 0;
  gfree = min(gfree, ngroups);
  // ---- tracking-free groups ----
#pragma unroll 1
  for (int g = 0; g < gfree; g++) {
#pragma unroll
    for (int ph = 0; ph < 4; ph++) {
      const int k = 4 * g + ph;
      const uint32_t selA = v | prevsh;
      prevsh = v * 256u + 0x4400u;
      uint32_t mA = 0;
      if (NMASK) { const uint32_t ncm = ncol_mask16(v); mA = ncm | (ncm_prev << 16); ncm_prev = ncm; }
      uint32_t upA = prmt(he, h8A, 0x5410);
      uint32_t eA = prmt(he, eoA, 0x5432);
      uint32_t upB = prmt(h8A, h8B, 0x5432);
      uint32_t eB = prmt(eoA, eoB, 0x5432);
      he = HE[(k + 2) * kFastThreads];
      v = V[(k + 2) * kFastThreads];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        {
          uint32_t s2 = prmt(rowA[r], rowB[r], selA);
          if (NMASK) s2 = mask_sel(s2, mA, nsc2);
          const uint32_t d = __vadd2(pA[r], s2);
          const uint32_t dm = __vadd2(d, neg_oe2);
          const uint32_t h = __vimax3_s16x2_relu(d, fA[r], eA);
          fA[r] = __viaddmax_s16x2(fA[r], neg_ext2, dm);
          eA = __viaddmax_s16x2(eA, neg_ext2, dm);
          pA[r] = upA; upA = h;
        }
        {
          uint32_t s2 = prmt(rowC[r], rowD[r], selB);
          if (NMASK) s2 = mask_sel(s2, mB, nsc2);
          const uint32_t d = __vadd2(pB[r], s2);
          const uint32_t dm = __vadd2(d, neg_oe2);
          const uint32_t h = __vimax3_s16x2_relu(d, fB[r], eB);
          fB[r] = __viaddmax_s16x2(fB[r], neg_ext2, dm);
          eB = __viaddmax_s16x2(eB, neg_ext2, dm);
          pB[r] = upB; upB = h;
        }
      }
      HE[(k - 2) * kFastThreads] = prmt(upB, eB, 0x7632);
      h8A = upA; eoA = eA; h8B = upB; eoB = eB;
      selB = selA_prev; selA_prev = selA;
      if (NMASK) { mB = mA_prev; mA_prev = mA; }
    }
  }
  // ---- tracked groups (the pipelined keys start as zeros: never an improvement) ----
#pragma unroll 1
  for (int g = gfree; g < ngroups; g++) {
#pragma unroll
    for (int ph = 0; ph < 4; ph++) {
      const int k = 4 * g + ph;
      const uint32_t selA = v | prevsh;
      prevsh = v * 256u + 0x4400u;
      uint32_t mA = 0;
      if (NMASK) { const uint32_t ncm = ncol_mask16(v); mA = ncm | (ncm_prev << 16); ncm_prev = ncm; }
      uint32_t upA = prmt(he, h8A, 0x5410);
      uint32_t eA = prmt(he, eoA, 0x5432);
      uint32_t upB = prmt(h8A, h8B, 0x5432);
      uint32_t eB = prmt(eoA, eoB, 0x5432);
      he = HE[(k + 2) * kFastThreads];
      v = V[(k + 2) * kFastThreads];
      // fold the PREVIOUS step's keys (software-pipelined: no input from this step); a group is closed when its phase-3 keys are in
      gA = key_tree9(gA, keyA);
      gB = key_tree9(gB, keyB);
      if (ph == 0) {
        const uint32_t pos = pos_base + (uint32_t)g - one;      // the group that just closed (first tracked group: nothing but the zero keys)
        track_update(lb.a, gA | 0x001F001Fu, gA, pos, one);
        track_update(lb.b, gB | 0x001F001Fu, gB, pos, one);
        gA = 0; gB = 0;
      }
#pragma unroll
      for (int r = 0; r < 8; r++) {
        {
          uint32_t s2 = prmt(rowA[r], rowB[r], selA);
          if (NMASK) s2 = mask_sel(s2, mA, nsc2);
          const uint32_t d = __vadd2(pA[r], s2);
          const uint32_t dm = __vadd2(d, neg_oe2);
          const uint32_t h = __vimax3_s16x2_relu(d, fA[r], eA);
          fA[r] = __viaddmax_s16x2(fA[r], neg_ext2, dm);
          eA = __viaddmax_s16x2(eA, neg_ext2, dm);
          keyA[r] = h * thirtytwo + (uint32_t)(((3 - ph) * 8 + (7 - r)) * 0x00010001);
          pA[r] = upA; upA = h;
        }
        {
          uint32_t s2 = prmt(rowC[r], rowD[r], selB);
          if (NMASK) s2 = mask_sel(s2, mB, nsc2);
          const uint32_t d = __vadd2(pB[r], s2);
          const uint32_t dm = __vadd2(d, neg_oe2);
          const uint32_t h = __vimax3_s16x2_relu(d, fB[r], eB);
          fB[r] = __viaddmax_s16x2(fB[r], neg_ext2, dm);
          eB = __viaddmax_s16x2(eB, neg_ext2, dm);
          keyB[r] = h * thirtytwo + (uint32_t)(((3 - ph) * 8 + (7 - r)) * 0x00010001);
          pB[r] = upB; upB = h;
        }
      }
      // B.H finished column k-3 (slot k-2): park (H,E) for the next quad-strip
      HE[(k - 2) * kFastThreads] = prmt(upB, eB, 0x7632);
      h8A = upA; eoA = eA; h8B = upB; eoB = eB;
      selB = selA_prev; selA_prev = selA;
      if (NMASK) { mB = mA_prev; mA_prev = mA; }
    }
  }
  if (gfree < ngroups) {
    gA = key_tree9(gA, keyA);
    gB = key_tree9(gB, keyB);
    const uint32_t pos = pos_base + (uint32_t)ngroups - one;
    track_update(lb.a, gA | 0x001F001Fu, gA, pos, one);
    track_update(lb.b, gB | 0x001F001Fu, gB, pos, one);
  }
}

__device__ __forceinline__ void lanebest_init(LaneBest& lb) { lb.bestK7 = 0x00070007u; lb.keyL = lb.keyH = 0; lb.posL = lb.posH = 0; }
__device__ __forceinline__ int lanebest_score_g4(const LaneBest& a, const LaneBest& b) {      // best score over the four lanes (keys score * 32 + 31)
  const int m = max(max((int)(short)(a.bestK7 & 0xFFFFu), (int)(short)(a.bestK7 >> 16)), max((int)(short)(b.bestK7 & 0xFFFFu), (int)(short)(b.bestK7 >> 16)));
  return m >> 5;
}
__device__ __forceinline__ void lanebest_init_g4(LaneBest& lb) { lb.bestK7 = 0x001F001Fu; lb.keyL = lb.keyH = 0; lb.posL = lb.posH = 0; }

// winner over the four records; pos = (quad-strip index << 16) | step
template <bool G4 = false>
__device__ __forceinline__ Winner pick_winner4(const LaneBest4& lb) {
  constexpr int SH = G4 ? 5 : 3;
  int sc[4], strip[4], col[4], row[4];
  sc[0] = (int)(short)(lb.a.bestK7 & 0xFFFFu) >> SH; sc[1] = (int)(short)(lb.a.bestK7 >> 16) >> SH;
  sc[2] = (int)(short)(lb.b.bestK7 & 0xFFFFu) >> SH; sc[3] = (int)(short)(lb.b.bestK7 >> 16) >> SH;
  const uint32_t pos[4] = {lb.a.posL, lb.a.posH, lb.b.posL, lb.b.posH};
  const uint32_t key[4] = {lb.a.keyL & 0xFFFFu, lb.a.keyH >> 16, lb.b.keyL & 0xFFFFu, lb.b.keyH >> 16};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    strip[i] = 4 * (int)(pos[i] >> 16) + i;
    // G4: pos holds the group of four steps, the key holds the step's phase inside it
    const int step = G4 ? 4 * (int)(pos[i] & 0xFFFFu) + (3 - (int)((key[i] >> 3) & 3u)) : (int)(pos[i] & 0xFFFFu);
    col[i] = step - i;
    row[i] = strip[i] * 8 + (7 - (int)(key[i] & 7u));
  }
  Winner w; w.score = 0; w.col = 0; w.row = 0;
  int bs = 0, bstrip = 0x7FFFFFFF;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (sc[i] > bs || (sc[i] == bs && sc[i] > 0 && strip[i] < bstrip)) { bs = sc[i]; bstrip = strip[i]; w.score = sc[i]; w.col = col[i]; w.row = row[i]; }
  }
  return w;
}

// ---- direction matrix from the packed lanes (LOCAL WITH_TB fast path) ---------------------------------------------------
// The four predicates of a direction nibble (local_kernel_template.h:45-60) are read off SIGN BITS of packed differences that
// cost one VIADD.16x2 each on the FMA pipe:
//   bit3 (F extended)  <=>  D-oe <= F-e   <=>  sign((D-oe-1) - (F-e))
//   bit2 (E extended)  <=>  D-oe <= E-e   <=>  sign((D-oe-1) - (E-e))
//   bit1               <=>  H != D        <=>  sign(D - H)                     (H >= D always)
//   bit0               <=>  H == D ? S < 0 : H == F   =  (~a & sign(S)) | (a & ~sign(F - H)), one LOP3 on the sign words
// Two PRMTs gather the sign-carrying bytes of the four words lane by lane, and one multiply by 0x00204081 moves the four sign
// bits of a lane into one nibble (the "movemask" multiplication), which a second multiply-add appends to the row word:
// 9 ALU + 11 FMA-pipe instructions per cell pair on top of the recurrence, no compares, no selects.
// (lop3_code0 / tb_nibbles live in dp_common.cuh: the wavefront traceback kernel uses them too)

// One quad-strip sweep with direction output.  tb points at this thread's element of tile 0 of strip 4j (A.L); tiles of one
// strip are 2*QR apart; `ts` = tb_stride.  valid[i]: strip 4j+i exists.  Each lane keeps the words of its last four columns
// in a 4-deep shift register and stores them as one uint4 when its column index reaches 3 mod 4; the lanes are one column
// apart, so exactly one of the four stores fires per step.  (A 4x unrolled variant with compile-time phases was 40 KB of
// code and ran at the speed of the int32 kernel: instruction-cache bound.)

__device__ __forceinline__ void tb_flush(uint4* __restrict__ tb, size_t ts, int QR2, int k, int ncol,
                                         uint32_t (&dAL)[4], uint32_t (&dAH)[4], uint32_t (&dBL)[4], uint32_t (&dBH)[4],
                                         uint32_t wAL, uint32_t wAH, uint32_t wBL, uint32_t wBH, const bool (&valid)[4]) {
  tb_push(dAL, wAL); tb_push(dAH, wAH); tb_push(dBL, wBL); tb_push(dBH, wBH);
  const int ph = k & 3;
  if (ph == 3) { const int c = k;     if (valid[0] && c < ncol)           tb[(size_t)(0 * QR2 + (c >> 2)) * ts] = make_uint4(dAL[0], dAL[1], dAL[2], dAL[3]); }
  if (ph == 0) { const int c = k - 1; if (valid[1] && c >= 3 && c < ncol) tb[(size_t)(1 * QR2 + (c >> 2)) * ts] = make_uint4(dAH[0], dAH[1], dAH[2], dAH[3]); }
  if (ph == 1) { const int c = k - 2; if (valid[2] && c >= 3 && c < ncol) tb[(size_t)(2 * QR2 + (c >> 2)) * ts] = make_uint4(dBL[0], dBL[1], dBL[2], dBL[3]); }
  if (ph == 2) { const int c = k - 3; if (valid[3] && c >= 3 && c < ncol) tb[(size_t)(3 * QR2 + (c >> 2)) * ts] = make_uint4(dBH[0], dBH[1], dBH[2], dBH[3]); }
}

template <bool NMASK = false>
__device__ __forceinline__ void local_sweep4_tb(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, int ncol,
                                                const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                                const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                                uint32_t neg_oe2, uint32_t neg_ext2, uint32_t eight, uint32_t one, uint32_t mm, uint32_t sixteen, uint32_t pos_base,
                                                LaneBest4& lb, uint4* __restrict__ tb, size_t ts, int QR2, const bool (&valid)[4], uint32_t nsc2 = 0) {
  uint32_t ncm_prev = 0, mB = 0, mA_prev = 0;
  uint32_t pA[8], fA[8], pB[8], fB[8], keyA[8], keyB[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { pA[r] = 0; fA[r] = 0; pB[r] = 0; fB[r] = 0; keyA[r] = 0; keyB[r] = 0; }
  uint32_t dAL[4] = {0, 0, 0, 0}, dAH[4] = {0, 0, 0, 0}, dBL[4] = {0, 0, 0, 0}, dBH[4] = {0, 0, 0, 0};
  uint32_t h8A = 0, eoA = 0, h8B = 0, eoB = 0;
  uint32_t prevsh = 0xCC00u;
  uint32_t selB = 0xCCCCu, selA_prev = 0xCCCCu;
  uint32_t he = HE[kFastThreads];
  uint32_t v = V[kFastThreads];
  uint32_t pos = pos_base;
  const uint32_t minus1_2 = 0xFFFFFFFFu;
  const int nsteps = ncol + 2 + (int)one;          // run-time 1: opaque trip count (see the compiler note in DESIGN.md)
GB200_TB_UNROLL_PRAGMA
  for (int k = 0; k < nsteps; k++) {
    const uint32_t selA = v | prevsh;
    prevsh = v * 256u + 0x4400u;
    uint32_t mA = 0;
    if (NMASK) { const uint32_t ncm = ncol_mask16(v); mA = ncm | (ncm_prev << 16); ncm_prev = ncm; }
    uint32_t upA = prmt(he, h8A, 0x5410);
    uint32_t eA = prmt(he, eoA, 0x5432);
    uint32_t upB = prmt(h8A, h8B, 0x5432);
    uint32_t eB = prmt(eoA, eoB, 0x5432);
    he = HE[(k + 2) * kFastThreads];
    v = V[(k + 2) * kFastThreads];
    {
      const uint32_t cm = key_tree(keyA);
      track_update(lb.a, cm | 0x00070007u, cm, pos, one);
      const uint32_t cn = key_tree(keyB);
      track_update(lb.b, cn | 0x00070007u, cn, pos, one);
      pos = pos_base + (uint32_t)k;
    }
    uint32_t wAL = 0, wAH = 0, wBL = 0, wBH = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      {
        uint32_t s2 = prmt(rowA[r], rowB[r], selA);
        if (NMASK) s2 = mask_sel(s2, mA, nsc2);
        const uint32_t d = __vadd2(pA[r], s2);
        const uint32_t dm = __vadd2(d, neg_oe2);
        const uint32_t h = __vimax3_s16x2_relu(d, fA[r], eA);
        tb_nibbles(d, dm, h, fA[r], eA, s2, neg_ext2, minus1_2, mm, sixteen, wAL, wAH);
        fA[r] = __viaddmax_s16x2(fA[r], neg_ext2, dm);
        eA = __viaddmax_s16x2(eA, neg_ext2, dm);
        keyA[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
        pA[r] = upA; upA = h;
      }
      {
        uint32_t s2 = prmt(rowC[r], rowD[r], selB);
        if (NMASK) s2 = mask_sel(s2, mB, nsc2);
        const uint32_t d = __vadd2(pB[r], s2);
        const uint32_t dm = __vadd2(d, neg_oe2);
        const uint32_t h = __vimax3_s16x2_relu(d, fB[r], eB);
        tb_nibbles(d, dm, h, fB[r], eB, s2, neg_ext2, minus1_2, mm, sixteen, wBL, wBH);
        fB[r] = __viaddmax_s16x2(fB[r], neg_ext2, dm);
        eB = __viaddmax_s16x2(eB, neg_ext2, dm);
        keyB[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
        pB[r] = upB; upB = h;
      }
    }
    HE[(k - 2) * kFastThreads] = prmt(upB, eB, 0x7632);
    h8A = upA; eoA = eA; h8B = upB; eoB = eB;
    selB = selA_prev; selA_prev = selA;
    if (NMASK) { mB = mA_prev; mA_prev = mA; }
    tb_flush(tb, ts, QR2, k, ncol, dAL, dAH, dBL, dBH, wAL, wAH, wBL, wBH, valid);
  }
  {
    const uint32_t cm = key_tree(keyA);
    track_update(lb.a, cm | 0x00070007u, cm, pos, one);
    const uint32_t cn = key_tree(keyB);
    track_update(lb.b, cn | 0x00070007u, cn, pos, one);
  }
}

// ---- BANDED + WITH_TB on the packed lanes ------------------------------------------------------------------------------------
// The engine's banded traceback (an extension: the reference's banded kernel has no traceback) is the LOCAL recurrence and its
// direction nibbles over the reference's TILE band (banded.h:35,83-85): strip i covers query words [lo(i), hi(i)), cells outside
// read as 0, every strip starts from a zero state at its first word (dp_generic.cuh banded_tb_forward).  A quad-strip sweep only
// walks the union of its four strips' bands.  Each lane has a window of steps in which it is inside ITS OWN band:
//   * at the step a lane enters its window, its halves of the diagonal / F registers are cleared (the strip's zero start state);
//   * whatever a lane hands on while it is outside its window (bottom H / E to the next lane or to shared memory) is masked to 0,
//     which is exactly what a cell outside the band reads as; its keys are masked out of the arg-max; its tiles are not stored.
// The sweep is cut into segments at the (at most eight) enter / exit steps, so the step body only carries six AND masks.
struct BandWin { int enter[4], exitk[4]; long long tile_base[4]; };

template <bool NMASK = false>
__device__ __forceinline__ void local_sweep4_tbb(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, const BandWin& bw,
                                                 const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                                 const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                                 uint32_t neg_oe2, uint32_t neg_ext2, uint32_t eight, uint32_t one, uint32_t mm, uint32_t sixteen, uint32_t pos_base,
                                                 LaneBest4& lb, uint4* __restrict__ tb, size_t ts, int c0, uint32_t nsc2 = 0) {
  uint32_t ncm_prev = 0, mB = 0, mA_prev = 0;
  uint32_t pA[8], fA[8], pB[8], fB[8], keyA[8], keyB[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { pA[r] = 0; fA[r] = 0; pB[r] = 0; fB[r] = 0; keyA[r] = 0; keyB[r] = 0; }
  uint32_t dAL[4] = {0, 0, 0, 0}, dAH[4] = {0, 0, 0, 0}, dBL[4] = {0, 0, 0, 0}, dBH[4] = {0, 0, 0, 0};
  uint32_t h8A = 0, eoA = 0, h8B = 0, eoB = 0;
  uint32_t prevsh = 0xCC00u;
  uint32_t selB = 0xCCCCu, selA_prev = 0xCCCCu;
  uint32_t he = HE[kFastThreads];
  uint32_t v = V[kFastThreads];
  uint32_t pos = pos_base;
  const uint32_t minus1_2 = 0xFFFFFFFFu;
  int nsteps = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) nsteps = max(nsteps, bw.exitk[i]);
  uint32_t inA = 0, inB = 0, inA_prev = 0, inB_prev = 0;
  int k = 0;
  while (k < nsteps) {
    // ---- segment boundary: lanes entering at k start from the zero state; masks of the lanes inside during [k, next)
    int next = nsteps;
    uint32_t keepA = 0xFFFFFFFFu, keepB = 0xFFFFFFFFu;
    inA = 0; inB = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (bw.enter[i] > k) next = min(next, bw.enter[i]);
      if (bw.exitk[i] > k) next = min(next, bw.exitk[i]);
      const bool inside = bw.enter[i] <= k && k < bw.exitk[i];
      const uint32_t half = (i & 1) ? 0xFFFF0000u : 0x0000FFFFu;
      if (inside) { if (i < 2) inA |= half; else inB |= half; }
      if (bw.enter[i] == k) { if (i < 2) keepA &= ~half; else keepB &= ~half; }
    }
    if (keepA != 0xFFFFFFFFu) {
#pragma unroll
      for (int r = 0; r < 8; r++) { pA[r] &= keepA; fA[r] &= keepA; }
    }
    if (keepB != 0xFFFFFFFFu) {
#pragma unroll
      for (int r = 0; r < 8; r++) { pB[r] &= keepB; fB[r] &= keepB; }
    }
    const bool stA0 = (inA & 0xFFFFu) != 0, stA1 = (inA >> 16) != 0, stB0 = (inB & 0xFFFFu) != 0, stB1 = (inB >> 16) != 0;
GB200_TBB_UNROLL_PRAGMA
    for (; k < next; k++) {
      const uint32_t selA = v | prevsh;
      prevsh = v * 256u + 0x4400u;
      uint32_t mA = 0;
      if (NMASK) { const uint32_t ncm = ncol_mask16(v); mA = ncm | (ncm_prev << 16); ncm_prev = ncm; }
      uint32_t upA = prmt(he, h8A, 0x5410);
      uint32_t eA = prmt(he, eoA, 0x5432);
      uint32_t upB = prmt(h8A, h8B, 0x5432);
      uint32_t eB = prmt(eoA, eoB, 0x5432);
      he = HE[(k + 2) * kFastThreads];
      v = V[(k + 2) * kFastThreads];
      {
        const uint32_t cm = key_tree(keyA) & inA_prev;               // the keys are the previous step's: masked with that step's windows
        track_update(lb.a, cm | 0x00070007u, cm, pos, one);
        const uint32_t cn = key_tree(keyB) & inB_prev;
        track_update(lb.b, cn | 0x00070007u, cn, pos, one);
        pos = pos_base + (uint32_t)k;
      }
      uint32_t wAL = 0, wAH = 0, wBL = 0, wBH = 0;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        {
          uint32_t s2 = prmt(rowA[r], rowB[r], selA);
          if (NMASK) s2 = mask_sel(s2, mA, nsc2);
          const uint32_t d = __vadd2(pA[r], s2);
          const uint32_t dm = __vadd2(d, neg_oe2);
          const uint32_t h = __vimax3_s16x2_relu(d, fA[r], eA);
          tb_nibbles(d, dm, h, fA[r], eA, s2, neg_ext2, minus1_2, mm, sixteen, wAL, wAH);
          fA[r] = __viaddmax_s16x2(fA[r], neg_ext2, dm);
          eA = __viaddmax_s16x2(eA, neg_ext2, dm);
          keyA[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
          pA[r] = upA; upA = h;
        }
        {
          uint32_t s2 = prmt(rowC[r], rowD[r], selB);
          if (NMASK) s2 = mask_sel(s2, mB, nsc2);
          const uint32_t d = __vadd2(pB[r], s2);
          const uint32_t dm = __vadd2(d, neg_oe2);
          const uint32_t h = __vimax3_s16x2_relu(d, fB[r], eB);
          tb_nibbles(d, dm, h, fB[r], eB, s2, neg_ext2, minus1_2, mm, sixteen, wBL, wBH);
          fB[r] = __viaddmax_s16x2(fB[r], neg_ext2, dm);
          eB = __viaddmax_s16x2(eB, neg_ext2, dm);
          keyB[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
          pB[r] = upB; upB = h;
        }
      }
      // what a lane outside its band hands on reads as 0
      upA &= inA; eA &= inA; upB &= inB; eB &= inB;
      HE[(k - 2) * kFastThreads] = prmt(upB, eB, 0x7632);
      h8A = upA; eoA = eA; h8B = upB; eoB = eB;
      selB = selA_prev; selA_prev = selA;
      inA_prev = inA; inB_prev = inB;
      if (NMASK) { mB = mA_prev; mA_prev = mA; }
      // one of the four lanes completes a tile of four columns per step (its column index reaches 3 mod 4)
      tb_push(dAL, wAL); tb_push(dAH, wAH); tb_push(dBL, wBL); tb_push(dBH, wBH);
      const int ph = k & 3;
      if (ph == 3) { if (stA0) tb[(size_t)(bw.tile_base[0] + ((c0 + k) >> 2)) * ts] = make_uint4(dAL[0], dAL[1], dAL[2], dAL[3]); }
      if (ph == 0) { if (stA1) tb[(size_t)(bw.tile_base[1] + ((c0 + k - 1) >> 2)) * ts] = make_uint4(dAH[0], dAH[1], dAH[2], dAH[3]); }
      if (ph == 1) { if (stB0) tb[(size_t)(bw.tile_base[2] + ((c0 + k - 2) >> 2)) * ts] = make_uint4(dBL[0], dBL[1], dBL[2], dBL[3]); }
      if (ph == 2) { if (stB1) tb[(size_t)(bw.tile_base[3] + ((c0 + k - 3) >> 2)) * ts] = make_uint4(dBH[0], dBH[1], dBH[2], dBH[3]); }
    }
  }
  {
    const uint32_t cm = key_tree(keyA) & inA_prev;
    track_update(lb.a, cm | 0x00070007u, cm, pos, one);
    const uint32_t cn = key_tree(keyB) & inB_prev;
    track_update(lb.b, cn | 0x00070007u, cn, pos, one);
  }
}

// ---- score-only BANDED (banded.h:10-139) on the same window machinery: local-type with the zero clamp, but H-BASED gap opening
// (F from H of the previous column, E from H of the row above: banded.h:99-114), so it is not the LOCAL recurrence of the traceback
// sweep above.  Chain formulation per cell pair: PRMT (substitution), VIADDMNMX (F), VIADDMNMX (E), VIMNMX3.RELU (H) on the ALU pipe;
// D = diag + s, H - oe (feeds the next row's E and the next column's F) and the arg-max key on the FMA pipe.  The parked / handed-on
// E is the row's OWN E (the next row advances it), and (0, 0) for whatever lies outside a strip's band: every H is clamped at 0, so
// F / E values <= 0 are interchangeable and the cleared start state (F = 0, H - oe = 0) of a lane that enters its window is exact.
// (ptxas folds the single-use add D = diag + s into the H maximum: VIADDMNMX + VIMNMX on the ALU pipe instead of VIADD.16x2 +
// VIMNMX3, 6.9 instead of 5.9 ALU-pipe instructions per cell pair.  Giving D a second use through the keys was measured: keys built
// from D need non-negative lanes (the packed multiply carries a negative low lane into the high one), and D times a run-time zero
// costs the IMAD it saves: 3.17 ms vs 3.15 ms on config 4.  Kept simple.)
template <bool NMASK = false>
__device__ __forceinline__ void banded_sweep4(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, const BandWin& bw,
                                              const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                              const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                              uint32_t neg_oe2, uint32_t neg_ext2, uint32_t eight, uint32_t one, uint32_t pos_base,
                                              LaneBest4& lb, uint32_t nsc2 = 0) {
  uint32_t ncm_prev = 0, mB = 0, mA_prev = 0;
  uint32_t pA[8], fA[8], glA[8], pB[8], fB[8], glB[8], keyA[8], keyB[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { pA[r] = 0; fA[r] = 0; glA[r] = 0; pB[r] = 0; fB[r] = 0; glB[r] = 0; keyA[r] = 0; keyB[r] = 0; }
  uint32_t h8A = 0, eoA = 0, h8B = 0, eoB = 0;
  uint32_t prevsh = 0xCC00u;
  uint32_t selB = 0xCCCCu, selA_prev = 0xCCCCu;
  uint32_t he = HE[kFastThreads];
  uint32_t v = V[kFastThreads];
  uint32_t pos = pos_base;
  int nsteps = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) nsteps = max(nsteps, bw.exitk[i]);
  uint32_t inA = 0, inB = 0, inA_prev = 0, inB_prev = 0;
  int k = 0;
  while (k < nsteps) {
    int next = nsteps;
    uint32_t keepA = 0xFFFFFFFFu, keepB = 0xFFFFFFFFu;
    inA = 0; inB = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (bw.enter[i] > k) next = min(next, bw.enter[i]);
      if (bw.exitk[i] > k) next = min(next, bw.exitk[i]);
      const bool inside = bw.enter[i] <= k && k < bw.exitk[i];
      const uint32_t half = (i & 1) ? 0xFFFF0000u : 0x0000FFFFu;
      if (inside) { if (i < 2) inA |= half; else inB |= half; }
      if (bw.enter[i] == k) { if (i < 2) keepA &= ~half; else keepB &= ~half; }
    }
    if (keepA != 0xFFFFFFFFu) {
#pragma unroll
      for (int r = 0; r < 8; r++) { pA[r] &= keepA; fA[r] &= keepA; glA[r] &= keepA; }
    }
    if (keepB != 0xFFFFFFFFu) {
#pragma unroll
      for (int r = 0; r < 8; r++) { pB[r] &= keepB; fB[r] &= keepB; glB[r] &= keepB; }
    }
GB200_BSC_UNROLL_PRAGMA
    for (; k < next; k++) {
      const uint32_t selA = v | prevsh;
      prevsh = v * 256u + 0x4400u;
      uint32_t mA = 0;
      if (NMASK) { const uint32_t ncm = ncol_mask16(v); mA = ncm | (ncm_prev << 16); ncm_prev = ncm; }
      uint32_t upA = prmt(he, h8A, 0x5410);
      uint32_t eA = prmt(he, eoA, 0x5432);
      uint32_t upB = prmt(h8A, h8B, 0x5432);
      uint32_t eB = prmt(eoA, eoB, 0x5432);
      he = HE[(k + 2) * kFastThreads];
      v = V[(k + 2) * kFastThreads];
      {
        const uint32_t cm = key_tree(keyA) & inA_prev;               // the keys are the previous step's: masked with that step's windows
        track_update(lb.a, cm | 0x00070007u, cm, pos, one);
        const uint32_t cn = key_tree(keyB) & inB_prev;
        track_update(lb.b, cn | 0x00070007u, cn, pos, one);
        pos = pos_base + (uint32_t)k;
      }
      uint32_t gupA = __vadd2(upA, neg_oe2), gupB = __vadd2(upB, neg_oe2);       // H of the row above - oe
#pragma unroll
      for (int r = 0; r < 8; r++) {
        {
          uint32_t s2 = prmt(rowA[r], rowB[r], selA);
          if (NMASK) s2 = mask_sel(s2, mA, nsc2);
          const uint32_t d = __vadd2(pA[r], s2);
          fA[r] = __viaddmax_s16x2(fA[r], neg_ext2, glA[r]);          // F = max(F - e, H(left) - oe)      (banded.h:101)
          eA = __viaddmax_s16x2(eA, neg_ext2, gupA);                  // E = max(E - e, H(above) - oe)     (:102)
          const uint32_t h = __vimax3_s16x2_relu(d, fA[r], eA);
          gupA = __vadd2(h, neg_oe2); glA[r] = gupA;
          keyA[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
          pA[r] = upA; upA = h;
        }
        {
          uint32_t s2 = prmt(rowC[r], rowD[r], selB);
          if (NMASK) s2 = mask_sel(s2, mB, nsc2);
          const uint32_t d = __vadd2(pB[r], s2);
          fB[r] = __viaddmax_s16x2(fB[r], neg_ext2, glB[r]);
          eB = __viaddmax_s16x2(eB, neg_ext2, gupB);
          const uint32_t h = __vimax3_s16x2_relu(d, fB[r], eB);
          gupB = __vadd2(h, neg_oe2); glB[r] = gupB;
          keyB[r] = h * eight + (uint32_t)((7 - r) * 0x00010001);
          pB[r] = upB; upB = h;
        }
      }
      // what a lane outside its band hands on reads as 0
      upA &= inA; eA &= inA; upB &= inB; eB &= inB;
      HE[(k - 2) * kFastThreads] = prmt(upB, eB, 0x7632);
      h8A = upA; eoA = eA; h8B = upB; eoB = eB;
      selB = selA_prev; selA_prev = selA;
      inA_prev = inA; inB_prev = inB;
      if (NMASK) { mB = mA_prev; mA_prev = mA; }
    }
  }
  {
    const uint32_t cm = key_tree(keyA) & inA_prev;
    track_update(lb.a, cm | 0x00070007u, cm, pos, one);
    const uint32_t cn = key_tree(keyB) & inB_prev;
    track_update(lb.b, cn | 0x00070007u, cn, pos, one);
  }
}

// Does the target hold a base outside A/C/G/T inside its true length?  Word-parallel (a nibble is valid iff it is 0001, 0011,
// 0100 or 0111) and WITHOUT an early exit: the loads of a thread are independent and pipeline, where the per-nibble loop with
// `&& !odd` serialised 75 dependent global loads per pair (7 % of the local kernel's time in the ncu source view).
__device__ __forceinline__ uint32_t odd_nibbles(uint32_t w) {   // bit 4k set iff nibble k is not A/C/G/T
  const uint32_t b0 = w, b1 = w >> 1, b2 = w >> 2, b3 = w >> 3;
  const uint32_t valid = ~b3 & ((~b2 & b0) | (b2 & ~(b1 ^ b0)));
  return ~valid & 0x11111111u;
}
__device__ __forceinline__ bool target_odd_scan(const uint32_t* __restrict__ tw, uint32_t tlen, int TR) {
  (void)TR;
  uint32_t acc = 0;
  const int full = (int)(tlen >> 3), rem = (int)(tlen & 7u);
#pragma unroll 8
  for (int i = 0; i < full; i++) acc |= odd_nibbles(tw[i]);
  if (rem) acc |= odd_nibbles(tw[full]) & (0x11111111u << (4 * (8 - rem)));      // first base in bits 31:28: keep the top `rem` nibbles
  return acc != 0;
}

// N-aware pass: nibbles outside A/C/G/T that are not the N code either (IUPAC letters ...) -- only the general kernel scores those
__device__ __forceinline__ uint32_t exotic_nibbles(uint32_t w, uint32_t nval) {
  const uint32_t z = w ^ (nval * 0x11111111u);
  const uint32_t isn = ~(z | (z >> 1) | (z >> 2) | (z >> 3)) & 0x11111111u;
  return odd_nibbles(w) & ~isn;
}
__device__ __forceinline__ bool target_exotic_scan(const uint32_t* __restrict__ tw, uint32_t tlen, uint32_t nval) {
  uint32_t acc = 0;
  const int full = (int)(tlen >> 3), rem = (int)(tlen & 7u);
#pragma unroll 8
  for (int i = 0; i < full; i++) acc |= exotic_nibbles(tw[i], nval);
  if (rem) acc |= exotic_nibbles(tw[full], nval) & (0x11111111u << (4 * (8 - rem)));
  return acc != 0;
}
// row table of the N-aware pass: an N row scores n_score against every query base
__device__ __forceinline__ uint32_t row_table_n(uint32_t code, uint32_t mis4, uint32_t delta, uint32_t nval, uint32_t nsc4) {
  return (code == nval) ? nsc4 : row_table(base_idx(code), mis4, delta);
}
template <bool NMASK>
__device__ __forceinline__ void quad_tables(uint32_t wa, uint32_t wb, uint32_t wc, uint32_t wd, uint32_t mis4, uint32_t delta, uint32_t nval, uint32_t nsc4,
                                            uint32_t (&rowA)[8], uint32_t (&rowB)[8], uint32_t (&rowC)[8], uint32_t (&rowD)[8]) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    if (NMASK) {
      rowA[r] = row_table_n(nib(wa, r), mis4, delta, nval, nsc4); rowB[r] = row_table_n(nib(wb, r), mis4, delta, nval, nsc4);
      rowC[r] = row_table_n(nib(wc, r), mis4, delta, nval, nsc4); rowD[r] = row_table_n(nib(wd, r), mis4, delta, nval, nsc4);
    } else {
      rowA[r] = row_table(base_idx(nib(wa, r)), mis4, delta); rowB[r] = row_table(base_idx(nib(wb, r)), mis4, delta);
      rowC[r] = row_table(base_idx(nib(wc, r)), mis4, delta); rowD[r] = row_table(base_idx(nib(wd, r)), mis4, delta);
    }
  }
}
// selector byte of one query column.  Main pass: anything outside A/C/G/T flags the pair (`flag`).  N-aware pass: a real N gets the
// mask marker, pads stay benign, any other code flags the pair for the general kernel.
template <bool NMASK>
__device__ __forceinline__ uint8_t col_selector(uint32_t code, bool pad, uint32_t nval, bool& flag) {
  const uint32_t idx = base_idx(code);
  if (idx <= 3) return (uint8_t)(0x80u + 0x11u * idx);
  if (!NMASK) { flag |= !pad; return 0x88; }
  if (pad) return 0x88;
  if (code == nval) return (uint8_t)kSelN;
  flag = true; return 0x88;
}

// ---- k_classify: which pass scores a pair? --------------------------------------------------------------------------------
// The packed sequences are read once, EIGHT WORDS PER THREAD (thread t: pair t / chunks, words 8 (t % chunks) .. + 7 of the pair's
// query-then-target word sequence; neighbouring threads read neighbouring 32-byte pieces), and every pair that holds a base outside
// A/C/G/T inside its true lengths is flagged and listed: bit 0 = an N (N-aware launch, worklist), bit 1 = any other code or an
// empty sequence (general kernel, worklist2; wins over bit 0).  Because the lists exist BEFORE the main launch, the N-aware launch
// runs beside it on a second stream instead of behind it, so a batch with 1 % of N-carrying reads no longer pays a whole extra
// (nearly empty) wave, and the main kernel needs no target scan.  flags[] is zeroed by the launcher.
__global__ void __launch_bounds__(256) k_classify(const FastArgs fa, uint32_t* __restrict__ flags, uint32_t chunks, int nmask_ok, int zero_q_is_odd, int zero_t_is_odd) {
  const DpArgs& a = fa.a;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t pair = t / chunks, chunk = t - pair * chunks;
  if (pair >= (uint32_t)a.n) return;
  const uint32_t qlen = a.q_lens[pair], tlen = a.t_lens[pair];
  const uint32_t nq = n_words(qlen), nt = n_words(tlen), nval = a.n_value;
  uint32_t odd = 0, exo = 0;
  bool bX = (chunk == 0) && ((zero_q_is_odd && qlen == 0) || (zero_t_is_odd && tlen == 0));
  const uint32_t w0 = 8 * chunk;
  if (w0 < nq + nt) {
    const uint32_t* qw = a.pq + (a.q_offs[pair] >> 3);
    const uint32_t* tw = a.pt + (a.t_offs[pair] >> 3);
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {                                       // all loads first: eight independent requests per thread
      const uint32_t w = w0 + i;
      x[i] = (w < nq) ? qw[w] : ((w < nq + nt) ? tw[w - nq] : 0x11111111u);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t w = w0 + i;
      if (w >= nq + nt) continue;
      const bool isq = w < nq;
      const uint32_t left = (isq ? qlen : tlen) - 8 * (isq ? w : w - nq);  // >= 1 bases of this word are real
      const uint32_t m = left >= 8 ? 0x11111111u : (0x11111111u << (4 * (8 - left)));
      odd |= odd_nibbles(x[i]) & m; exo |= exotic_nibbles(x[i], nval) & m;
    }
  }
  bX |= (exo != 0);
  bool bN = (odd & ~exo) != 0;
  if (!nmask_ok) { bX |= bN; bN = false; }
  if (bN || bX) {
    const uint32_t bits = (bN ? 1u : 0u) | (bX ? 2u : 0u);
    const uint32_t old = atomicOr(flags + pair, bits);
    if ((bits & 2u) && !(old & 2u)) fa.worklist2[atomicAdd(fa.work_count + 1, 1u)] = pair;
    if ((bits & 1u) && !(old & 1u)) fa.worklist[atomicAdd(fa.work_count, 1u)] = pair;
  }
}

template <int START, bool NMASK = false, bool G4 = false>
__global__ void GB200_FAST_LB k_local_s16(const FastArgs fa) {
  extern __shared__ uint32_t smem[];
  const DpArgs& a = fa.a;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (NMASK) { if (tid >= fa.work_count[0]) return; tid = fa.worklist[tid]; if (fa.flags[tid] & 2u) return; }   // N-aware pass: one thread per listed pair
  else { if (tid >= (uint32_t)a.n) return; if (fa.flags && fa.flags[tid]) return; }   // flagged pairs belong to the other passes (k_classify)
  uint32_t* HE = smem + 2 * kFastThreads + threadIdx.x;          // slot s holds column s - 1; slots -2..0 are dummies
  uint8_t* V = (uint8_t*)(smem + (fa.q8max + kSlotPad) * kFastThreads) + threadIdx.x;

  const uint32_t* qw = a.pq + (a.q_offs[tid] >> 3);
  const uint32_t* tw = a.pt + (a.t_offs[tid] >> 3);
  const uint32_t qlen = a.q_lens[tid], tlen = a.t_lens[tid];
  const int QR = (int)n_words(qlen), TR = (int)n_words(tlen);
  const uint32_t mis4 = ((uint32_t)(-a.mismatch) & 0xFFu) * 0x01010101u;
  const uint32_t delta = ((uint32_t)a.match ^ (uint32_t)(-a.mismatch)) & 0xFFu;
  const uint32_t neg_oe2 = splat16(-(a.gap_open + a.gap_extend)), neg_ext2 = splat16(-a.gap_extend);
  const uint32_t nval = a.n_value, nsc2 = splat16(a.n_score), nsc4 = ((uint32_t)a.n_score & 0xFFu) * 0x01010101u;

  // ---- prologue: selector bytes of the query, (H,E) = 0; detect bases this pass cannot score ----
  bool odd = false;
  const int ncol = QR * 8;
  V[0] = 0x88; HE[0] = 0;
  for (int j = 0; j < QR; j++) {
    const uint32_t w = qw[j];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int c = 8 * j + k;
      V[(c + 1) * kFastThreads] = col_selector<NMASK>(nib(w, k), (uint32_t)c >= qlen, nval, odd);
      HE[(c + 1) * kFastThreads] = 0;
    }
  }
#pragma unroll
  for (int x = 1; x <= 5; x++) { V[(ncol + x) * kFastThreads] = 0x88; HE[(ncol + x) * kFastThreads] = 0; }
  if (odd) {   // cannot happen (k_classify filed this pair for this pass from the same words); kept as a safety net: general kernel
    const uint32_t slot = atomicAdd(fa.work_count + 1, 1u);
    fa.worklist2[slot] = tid;
    return;
  }

  // ---- forward pass over quad-strips ----
  const int mstep = NMASK ? max(a.match, a.n_score) : a.match;      // the most one diagonal step can add (score bounds of the prunings below)
  LaneBest4 lb;
  if (G4) { lanebest_init_g4(lb.a); lanebest_init_g4(lb.b); } else { lanebest_init(lb.a); lanebest_init(lb.b); }
  const int NQ = (TR + 3) >> 2;
  const int nsteps = ncol + 2 + (int)fa.one;   // run-time 1: keeps the trip count opaque (compiler note)
  if (START == START_TB) {
    // WITH_TB: same sweep, plus the 4-bit direction matrix in the reference's tile layout (local_kernel_template.h:257,380):
    // tile (strip * QR + word) * 2 + half of pair `tid` at tb[tile * tb_stride + tid]; get_tb walks it afterwards
    for (int j = 0; j < NQ; j++) {
      uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
      const bool valid[4] = {4 * j < TR, 4 * j + 1 < TR, 4 * j + 2 < TR, 4 * j + 3 < TR};
      const uint32_t wa = tw[4 * j];
      const uint32_t wb = valid[1] ? tw[4 * j + 1] : 0u;
      const uint32_t wc = valid[2] ? tw[4 * j + 2] : 0u;
      const uint32_t wd = valid[3] ? tw[4 * j + 3] : 0u;
      quad_tables<NMASK>(wa, wb, wc, wd, mis4, delta, nval, nsc4, rowA, rowB, rowC, rowD);
      uint4* tbq = a.tb + (size_t)(4 * j) * (size_t)(2 * QR) * a.tb_stride + tid;
      local_sweep4_tb<NMASK>(HE, V, ncol, rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, fa.eight, fa.one, fa.mm, fa.sixteen, (uint32_t)j << 16, lb,
                             tbq, (size_t)a.tb_stride, 2 * QR, valid, nsc2);
    }
    const Winner tw_ = pick_winner4(lb);
    a.score[tid] = tw_.score; a.q_end[tid] = tw_.col; a.t_end[tid] = tw_.row;
    return;
  }
  for (int j = 0; j < NQ; j++) {
    uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
    const uint32_t wa = tw[4 * j];
    const uint32_t wb = (4 * j + 1 < TR) ? tw[4 * j + 1] : 0u;     // phantom strips: every base mismatches
    const uint32_t wc = (4 * j + 2 < TR) ? tw[4 * j + 2] : 0u;
    const uint32_t wd = (4 * j + 3 < TR) ? tw[4 * j + 3] : 0u;
    quad_tables<NMASK>(wa, wb, wc, wd, mis4, delta, nval, nsc4, rowA, rowB, rowC, rowD);
    // G4: ncol / 4 + 1 groups = ncol + 4 steps (one phantom step more than the skew needs; its slots are dummies)
    // (columns q with match * (q + 1) <= best so far cannot improve it: their groups of steps run without the arg-max)
    // (the count is made WARP-UNIFORM -- the minimum over the converged lanes, which is conservative for each of them -- or the two
    // loops of the sweep would run divergent: measured 1.5x SLOWER with per-pair counts)
    if (G4) local_sweep4g<NMASK>(HE, V, (ncol >> 2) + (int)fa.one,
                                 GB200_PRUNE_ARGMAX && mstep > 0 ? (int)__reduce_min_sync(__activemask(), (unsigned)((lanebest_score_g4(lb.a, lb.b) / mstep) >> 2)) : 0,
                                 rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, fa.thirtytwo, fa.one, (uint32_t)j << 16, lb, nsc2);
    else local_sweep4<NMASK>(HE, V, nsteps, rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, fa.eight, fa.one, (uint32_t)j << 16, lb, nsc2);
  }
  const Winner fw = pick_winner4<G4>(lb);
  a.score[tid] = fw.score; a.q_end[tid] = fw.col; a.t_end[tid] = fw.row;

  if (START == START_POS) {
    // ---- reverse pass (local_kernel_template.h:441-511): same recurrence over the reversed words that hold the
    // end cell; the first cell that reaches the forward score is the start (target coordinate mirrored, Q3) ----
    int sx = 0, sy = 0;
    if (fw.score > 0) {
      const int rend_reg = min((fw.col >> 3) + 1, QR), gend_reg = min((fw.row >> 3) + 1, TR);
      const int rcol = rend_reg * 8;
      bool dummy = false;
      for (int j = 0; j < rend_reg; j++) {
        const uint32_t w = nibrev(qw[rend_reg - 1 - j]);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int c = 8 * j + k;
          V[(c + 1) * kFastThreads] = col_selector<NMASK>(nib(w, k), (uint32_t)(rcol - 1 - c) >= qlen, nval, dummy);
          HE[(c + 1) * kFastThreads] = 0;
        }
      }
#pragma unroll
      for (int x = 1; x <= 5; x++) { V[(rcol + x) * kFastThreads] = 0x88; HE[(rcol + x) * kFastThreads] = 0; }
      LaneBest4 rb;
      if (G4) { lanebest_init_g4(rb.a); lanebest_init_g4(rb.b); } else { lanebest_init(rb.a); lanebest_init(rb.b); }
      const int RD = (gend_reg + 3) >> 2;
      const int fwd7 = G4 ? ((fw.score << 5) | 31) : ((fw.score << 3) | 7);
      // Score-bound pruning of the reversed rectangle (exact).  A cell of this pass whose H reaches the forward score S is the start
      // of an optimal alignment that ENDS, in forward coordinates, on a cell with H = S inside the rectangle; the forward pass
      // recorded the FIRST such cell of the reference's (strip, column, row) order, so any other lies in the same strip at a column
      // >= the recorded one -- in this pass: every score-S path STARTS in the corner tile (row < 8, column < 8).  A path from
      // there through the cell (row i, column c):
      //   c - i = d > 0: at least d - 7 of its columns sit in horizontal gaps, which also cost it as many diagonal steps:
      //                  S <= m (rcol - (d - 7)) - open - (d - 7) ext        ->  d <= 7 + (m rcol - S - open) / (m + ext)
      //                  (and, whatever the gap costs, S <= m (rcol - d)     ->  d <= rcol - ceil(S / m))
      //   i - c = d > 0: at least d - 7 of its rows sit in vertical gaps:    S <= m rcol - open - (d - 7) ext  ->  d <= 7 + (m rcol - S - open) / ext
      // (m = the best substitution score; 7 alone when not even one gap is affordable).  Quad-strip j (rows 32 j .. 32 j + 31)
      // therefore only sweeps the columns [32 j - slackL, 32 (j + 1) + slackU): the cells outside read as the initial state (0, a
      // lower bound that the 0 clamp makes harmless), every cell on a path that reaches S keeps exact inputs and no cell exceeds
      // its true value, so the first cell that reaches S in the reference's order is the same.  For a 150 bp read with S = 130 the
      // five quad-strips sweep 47 + 70 + 70 + 70 + 47 columns instead of 5 x 152.
      // (both slacks are made warp-uniform, the maximum over the converged lanes: conservative for each of them, see the note on gfree)
      constexpr int kNoBound = 1 << 24;                               // more rows than any target has
      int slackU = kNoBound, slackL = kNoBound;
      if (GB200_PRUNE_REVERSE && mstep > 0) {
        const int num = mstep * rcol - fw.score - a.gap_open;
        const int up = min(rcol - (fw.score + mstep - 1) / mstep, 7 + (num >= 0 ? num / (mstep + a.gap_extend) : 0));
        const int lo = num < 0 ? 7 : (a.gap_extend > 0 ? 7 + num / a.gap_extend : kNoBound);     // free extension: a vertical gap of any length
        slackU = (int)__reduce_max_sync(__activemask(), (unsigned)max(up, 0));
        slackL = (int)__reduce_max_sync(__activemask(), (unsigned)min(lo, kNoBound));
      }
      const int gfree_abs = GB200_PRUNE_ARGMAX && mstep > 0 ? (int)__reduce_min_sync(__activemask(), (unsigned)(((fw.score - 1) / mstep) >> 2)) : 0;
      for (int j = 0; j < RD; j++) {
        const int c_lo = GB200_PRUNE_REVERSE_LOW ? min(max(32 * j - slackL, 0) & ~3, max(rcol - 4, 0)) : 0;
        const int lim = (int)min((long long)rcol, ((long long)32 * (j + 1) + slackU + 3) & ~3ll);     // > c_lo: rcol is a multiple of 8
        uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
        const uint32_t wa = nibrev(tw[gend_reg - 1 - 4 * j]);
        const uint32_t wb = (4 * j + 1 < gend_reg) ? nibrev(tw[gend_reg - 2 - 4 * j]) : 0u;
        const uint32_t wc = (4 * j + 2 < gend_reg) ? nibrev(tw[gend_reg - 3 - 4 * j]) : 0u;
        const uint32_t wd = (4 * j + 3 < gend_reg) ? nibrev(tw[gend_reg - 4 - 4 * j]) : 0u;
        quad_tables<NMASK>(wa, wb, wc, wd, mis4, delta, nval, nsc4, rowA, rowB, rowC, rowD);
        // (only cells that reach the forward score matter here: columns q with match * (q + 1) < fwd score need no arg-max)
        uint32_t* HEw = HE + (size_t)c_lo * kFastThreads;          // the sweep's column 0 is column c_lo: its zero start state IS the "reads as 0" left of the window
        const uint8_t* Vw = V + (size_t)c_lo * kFastThreads;
        if (G4) local_sweep4g<NMASK>(HEw, Vw, ((lim - c_lo) >> 2) + (int)fa.one, max(gfree_abs - (c_lo >> 2), 0),
                                     rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, fa.thirtytwo, fa.one, ((uint32_t)j << 16) + (uint32_t)(c_lo >> 2), rb, nsc2);
        else local_sweep4<NMASK>(HEw, Vw, (lim - c_lo) + 2 + (int)fa.one, rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, fa.eight, fa.one,
                                 ((uint32_t)j << 16) + (uint32_t)c_lo, rb, nsc2);
        if ((int)(short)(rb.a.bestK7 & 0xFFFFu) >= fwd7 || (int)(short)(rb.a.bestK7 >> 16) >= fwd7 ||
            (int)(short)(rb.b.bestK7 & 0xFFFFu) >= fwd7 || (int)(short)(rb.b.bestK7 >> 16) >= fwd7) break;
      }
      const Winner rw = pick_winner4<G4>(rb);
      if (rw.score > 0) {
        sx = rcol - 1 - rw.col;                                    // ridx
        const int strip = rw.row >> 3, r = rw.row & 7;
        sy = (8 * gend_reg - 1 - 8 * strip) + r;                   // gidx + (m-1): mirrored inside the word (Q3)
      }
    }
    a.q_start[tid] = sx; a.t_start[tid] = sy;
  }
}

// BANDED + WITH_TB: the LOCAL recurrence with direction nibbles over the tile band (see local_sweep4_tbb); only the band's tiles are
// stored: tile (strip * W + word - lo(strip)) * 2 + half, W = widest strip (dp_common.cuh band_geom), which k_get_tb addresses the same way
// TB = false: the reference's score-only BANDED kernel (banded.h: H-based opening, score + end cell) on the same tile band.
template <bool NMASK = false, bool TB = true>
__global__ void GB200_FAST_LB k_banded_tb_s16(const FastArgs fa) {
  extern __shared__ uint32_t smem[];
  const DpArgs& a = fa.a;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (NMASK) { if (tid >= fa.work_count[0]) return; tid = fa.worklist[tid]; if (fa.flags[tid] & 2u) return; }
  else { if (tid >= (uint32_t)a.n) return; if (fa.flags && fa.flags[tid]) return; }
  uint32_t* HE = smem + 2 * kFastThreads + threadIdx.x;          // slot s holds column s - 1; slots -2..0 are dummies
  uint8_t* V = (uint8_t*)(smem + (fa.q8max + kSlotPad) * kFastThreads) + threadIdx.x;
  const uint32_t* qw = a.pq + (a.q_offs[tid] >> 3);
  const uint32_t* tw = a.pt + (a.t_offs[tid] >> 3);
  const uint32_t qlen = a.q_lens[tid], tlen = a.t_lens[tid];
  const int QR = (int)n_words(qlen), TR = (int)n_words(tlen);
  const uint32_t mis4 = ((uint32_t)(-a.mismatch) & 0xFFu) * 0x01010101u;
  const uint32_t delta = ((uint32_t)a.match ^ (uint32_t)(-a.mismatch)) & 0xFFu;
  const uint32_t neg_oe2 = splat16(-(a.gap_open + a.gap_extend)), neg_ext2 = splat16(-a.gap_extend);
  const uint32_t nval = a.n_value, nsc2 = splat16(a.n_score), nsc4 = ((uint32_t)a.n_score & 0xFFu) * 0x01010101u;
  bool odd = false;
  const int ncol = QR * 8;
  V[0] = 0x88; HE[0] = 0;
  for (int j = 0; j < QR; j++) {
    const uint32_t w = qw[j];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int c = 8 * j + k;
      V[(c + 1) * kFastThreads] = col_selector<NMASK>(nib(w, k), (uint32_t)c >= qlen, nval, odd);
      HE[(c + 1) * kFastThreads] = 0;
    }
  }
#pragma unroll
  for (int x = 1; x <= 5; x++) { V[(ncol + x) * kFastThreads] = 0x88; HE[(ncol + x) * kFastThreads] = 0; }
  if (odd) { const uint32_t slot = atomicAdd(fa.work_count + 1, 1u); fa.worklist2[slot] = tid; return; }     // cannot happen (k_classify)

  const BandGeom bg = band_geom(QR, TR, a.k_band);
  LaneBest4 lb; lanebest_init(lb.a); lanebest_init(lb.b);
  const int NQ = (TR + 3) >> 2;
  for (int j = 0; j < NQ; j++) {
    uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
    BandWin bw;
    int ulo = 0x7FFFFFFF;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int strip = 4 * j + i;
      int lo = band_lo(bg, strip), hi = band_hi(bg, strip, QR);
      if (strip >= TR || lo >= hi) { lo = 0; hi = 0; }                // no cell of this strip is inside the band
      else ulo = min(ulo, lo);
      bw.enter[i] = lo; bw.exitk[i] = hi;                              // words for now
      bw.tile_base[i] = ((long long)strip * bg.W - lo) * 2;
    }
    if (ulo == 0x7FFFFFFF) continue;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool empty = bw.enter[i] >= bw.exitk[i];
      bw.enter[i] = empty ? 0 : (bw.enter[i] - ulo) * 8 + i;
      bw.exitk[i] = empty ? 0 : (bw.exitk[i] - ulo) * 8 + i;
    }
    const uint32_t wa = tw[4 * j];
    const uint32_t wb = (4 * j + 1 < TR) ? tw[4 * j + 1] : 0u;
    const uint32_t wc = (4 * j + 2 < TR) ? tw[4 * j + 2] : 0u;
    const uint32_t wd = (4 * j + 3 < TR) ? tw[4 * j + 3] : 0u;
    quad_tables<NMASK>(wa, wb, wc, wd, mis4, delta, nval, nsc4, rowA, rowB, rowC, rowD);
    const int c0 = ulo * 8;
    if (TB) local_sweep4_tbb<NMASK>(HE + (size_t)c0 * kFastThreads, V + (size_t)c0 * kFastThreads, bw, rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, fa.eight, fa.one, fa.mm, fa.sixteen,
                                    ((uint32_t)j << 16) + (uint32_t)c0, lb, a.tb + tid, (size_t)a.tb_stride, c0, nsc2);
    else banded_sweep4<NMASK>(HE + (size_t)c0 * kFastThreads, V + (size_t)c0 * kFastThreads, bw, rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, fa.eight, fa.one,
                              ((uint32_t)j << 16) + (uint32_t)c0, lb, nsc2);
  }
  const Winner w_ = pick_winner4(lb);
  a.score[tid] = w_.score; a.q_end[tid] = w_.col; a.t_end[tid] = w_.row;
}

// =====================================================================================================
// Semi-global (H-based opening, reference semiglobal_kernel_template.h:17-28) and global (D-based,
// no clamp, global.h:4-12) in the same two-lane skewed layout.  These recurrences have no zero clamp,
// so (a) the lane that starts one column late (lane H) gets its strip-start state re-imposed after the
// phantom step 0, and (b) "minus infinity" is a finite floor that a single `- gap_extend` cannot wrap
// in 16 bits; eligibility (fast_path_eligible) proves every real value stays far above it.
constexpr int kNegSafe = -16384;

__device__ __forceinline__ int lane_lo(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int lane_hi(uint32_t v) { return (int)(short)(v >> 16); }
__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

struct NcTrack {
  int best, y;          // semi-global: running (maxHH, maxXY_y)
  int stop;             // reverse pass: stop once best >= stop
  int cap;              // global: captured H(tlen-1, qlen-1)
  int tie = 0;          // 1: a later row with the SAME score replaces the record (nvbio's BestSink, nvb_strip.cuh); 0: strict '>' (GASAL2)
  int capture = 0;      // 1: the H-based (semi-global) recurrence captures (caplane, caprow) instead of scanning the column (nvbio GLOBAL)
  int last = 0;         // start pass: highest lane of the quad-strip whose strip the reference would still have processed (its scan ran)
};
// TAIL in {QUERY, BOTH} (semiglobal_kernel_template.h:185-216): the reference scans the bottom row of the LAST PROCESSED strip over the
// query columns, strict '>' left to right.  Which strip that is may only be known after the sweep (the start pass stops between
// strips), so the sweep keeps, for each of its four lanes, the first maximum of that lane's bottom row over the true query columns.
struct NcBottom {
  LaneBest a, b;              // running (value, step) per lane: a = lanes A.L / A.H, b = B.L / B.H; values un-biased
  uint32_t biasA, biasB;      // frame bias of the bottom rows at the current step: (column + 7) ext per lane
  int nq;                     // true query columns from the sweep's first column on
};
__device__ __forceinline__ void ncbottom_init(NcBottom& bt, int col0, int qlen, int ext) {
  bt.a.bestK7 = bt.b.bestK7 = 0x80008000u; bt.a.keyL = bt.a.keyH = bt.b.keyL = bt.b.keyH = 0; bt.a.posL = bt.a.posH = bt.b.posL = bt.b.posH = 0;
  bt.biasA = pack2((col0 + 7) * ext, (col0 + 6) * ext); bt.biasB = pack2((col0 + 5) * ext, (col0 + 4) * ext);
  bt.nq = qlen - col0;
}
// (value, column) of lane i after the sweep
__device__ __forceinline__ void ncbottom_get(const NcBottom& bt, int lane, int col0, int& v, int& col) {
  const LaneBest& lb = (lane & 2) ? bt.b : bt.a;
  v = (lane & 1) ? lane_hi(lb.bestK7) : lane_lo(lb.bestK7);
  col = col0 + (int)((lane & 1) ? lb.posH : lb.posL) - lane;
}

template <int ALGO, bool REV>
__device__ __forceinline__ void nc_sweep(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, int nsteps,
                                         const uint32_t (&rowA)[8], const uint32_t (&rowB)[8], const uint32_t (&pinit)[8],
                                         uint32_t neg_oe2, uint32_t neg_ext2,
                                         int kscanL, int kscanH, int rowbaseL, int rowbaseH, int tlen, int caprow, NcTrack& tr) {
  constexpr bool SEMI = (ALGO == ALGO_SEMI_GLOBAL);
  const uint32_t neg2 = splat16(kNegSafe);
  uint32_t p[8], f[8], gl[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { p[r] = pinit[r]; f[r] = neg2; gl[r] = neg_oe2; }   // left column: H = 0 (semi), F = -inf
  uint32_t h8 = 0, eo = 0;
  uint32_t prevsh = 0xCC00u;
  uint32_t he = HE[kFastThreads];
  uint32_t v = V[kFastThreads];
  for (int k = 0; k < nsteps; k++) {
    const uint32_t sel = v | prevsh;
    prevsh = v * 256u + 0x4400u;
    uint32_t up = prmt(he, h8, 0x5410);
    uint32_t e = prmt(he, eo, 0x5432);
    he = HE[(k + 2) * kFastThreads];
    v = V[(k + 2) * kFastThreads];
    uint32_t hh[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint32_t s2 = prmt(rowA[r], rowB[r], sel);
      const uint32_t d = __vadd2(p[r], s2);
      uint32_t h;
      if (SEMI) {
        // H-based opening without the serial H -> E -> H chain: with M = max(D, F) and H = max(M, E),
        // E(t+1,q) = max(H - oe, E - e) = max(max(D, F) - oe, E - e) because oe >= e.  `e` therefore enters a row
        // as E(t,q) and leaves it as E(t+1,q); the parked / handed-over E is the NEXT row's.
        f[r] = __viaddmax_s16x2(f[r], neg_ext2, gl[r]);        // F(t,q) = max(H(t,q-1) - oe, F(t,q-1) - e)
        h = __vimax3_s16x2(d, f[r], e);
        const uint32_t dm = __vadd2(d, neg_oe2), fm = __vadd2(f[r], neg_oe2), ee = __vadd2(e, neg_ext2);
        e = __vimax3_s16x2(dm, fm, ee);
        gl[r] = __vadd2(h, neg_oe2);
      } else {
        const uint32_t dm = __vadd2(d, neg_oe2);
        h = __vimax3_s16x2(d, f[r], e);
        f[r] = __viaddmax_s16x2(f[r], neg_ext2, dm);
        e = __viaddmax_s16x2(e, neg_ext2, dm);
      }
      hh[r] = h;
      p[r] = up; up = h;
    }
    HE[k * kFastThreads] = prmt(up, e, 0x7632);
    h8 = up; eo = e;
    if (k == 0) {
      // lane H just ran its phantom column -1: restore its strip-start state
#pragma unroll
      for (int r = 0; r < 8; r++) {
        p[r] = prmt(p[r], pinit[r], 0x7610);
        f[r] = prmt(f[r], neg2, 0x7610);
        if (SEMI) gl[r] = prmt(gl[r], neg_oe2, 0x7610);
      }
    }
    if (SEMI) {
      // TAIL == TARGET: the last true query column, rows inside the target, strict '>' top-down (:160-178)
      if (k == kscanL) {
#pragma unroll
        for (int r = 0; r < 8; r++) { const int hv = lane_lo(hh[r]); if (rowbaseL + r < tlen && hv > tr.best) { tr.best = hv; tr.y = rowbaseL + r; } }
      }
      if (k == kscanH && (!REV || tr.best < tr.stop)) {
#pragma unroll
        for (int r = 0; r < 8; r++) { const int hv = lane_hi(hh[r]); if (rowbaseH + r < tlen && hv > tr.best) { tr.best = hv; tr.y = rowbaseH + r; } }
      }
    } else {
      if (k == kscanL) {
#pragma unroll
        for (int r = 0; r < 8; r++) if (r == caprow) tr.cap = lane_lo(hh[r]);
      }
      if (k == kscanH) {
#pragma unroll
        for (int r = 0; r < 8; r++) if (r == caprow) tr.cap = lane_hi(hh[r]);
      }
    }
  }
}

// ---- quad-strip variant of nc_sweep (see local_sweep4): strips 4j..4j+3 as A.L, A.H, B.L, B.H, skewed by one column each.
// The rare per-step work (re-imposing the strip-start state of the late lanes after their phantom columns; the scan of
// the last true query column; the global capture) is gathered behind ONE compare per step (k == knext).
//   kq      : step at which lane A.L is on the scanned / captured column (qlen - 1); lane i is there at step kq + i
//   lane_lo/lane_hi of hhA / hhB give the four lanes' H values of the current step
#ifndef GB200_NC_SPLIT
#define GB200_NC_SPLIT 1          // 1: rare steps (phantom-column fix-ups, scan / capture) run in their own loops, the hot loop has no rare code
#endif
#ifndef GB200_NC_UNROLL_SEMI
#define GB200_NC_UNROLL_SEMI 2
#endif
constexpr int kNcUnrollSemi = GB200_NC_UNROLL_SEMI;
struct NcState {
  uint32_t pA[8], fA[8], glA[8], pB[8], fB[8], glB[8];
  uint32_t h8A, eoA, h8B, eoB, prevsh, selB, selA_prev, he, v;
  uint32_t ncm_prev, mB, mA_prev;                 // NMASK: N-column masks, rotated like the selectors (see local_sweep4)
};

// one step (one query column per lane) of the quad-strip sweep; RARE adds the per-step work that only a handful of steps need
template <int ALGO, bool REV, bool NMASK, bool RARE, bool BROW = false>
__device__ __forceinline__ void nc_step(NcState& st, int k, uint32_t* __restrict__ HE, const uint8_t* __restrict__ V,
                                        const uint32_t (&rowA)[8], const uint32_t (&rowB)[8], const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                        const uint32_t (&pinitA)[8], const uint32_t (&pinitB)[8], const uint32_t (&hpinitA)[8], const uint32_t (&hpinitB)[8],
                                        const uint32_t (&finitA)[8], const uint32_t (&finitB)[8],
                                        uint32_t neg_oe2, uint32_t neg2, uint32_t conv8, int bias_ext, uint32_t nsc2,
                                        int kq, int rowbase, int tlen, int caplane, int caprow, NcTrack& tr, int col0 = 0, NcBottom* bt = nullptr) {
  constexpr bool SEMI = (ALGO == ALGO_SEMI_GLOBAL);
  const uint32_t selA = st.v | st.prevsh;
  st.prevsh = st.v * 256u + 0x4400u;
  uint32_t mA = 0;
  if (NMASK) { const uint32_t ncm = ncol_mask16(st.v); mA = ncm | (st.ncm_prev << 16); st.ncm_prev = ncm; }
  uint32_t upA = prmt(st.he, st.h8A, 0x5410);
  uint32_t eA = prmt(st.he, st.eoA, 0x5432);
  uint32_t upB = prmt(st.h8A, st.h8B, 0x5432);
  uint32_t eB = prmt(st.eoA, st.eoB, 0x5432);
  st.he = HE[(k + 2) * kFastThreads];
  st.v = V[(k + 2) * kFastThreads];
  uint32_t hhA[8], hhB[8];
  // SEMI keeps H of the previous column in glA/glB anyway, and the diagonal input of row r is exactly that H of row r - 1:
  // reading it from there (pure renaming once the loop is unrolled) instead of a second rotating array saves 16 register moves per step
  uint32_t dgA[8], dgB[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { dgA[r] = SEMI ? (r ? st.glA[r - 1] : st.pA[0]) : st.pA[r]; dgB[r] = SEMI ? (r ? st.glB[r - 1] : st.pB[0]) : st.pB[r]; }
  if (SEMI) { st.pA[0] = upA; st.pB[0] = upB; }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    {
      uint32_t s2 = prmt(rowA[r], rowB[r], selA);
      if (NMASK) s2 = mask_sel(s2, mA, nsc2);
      const uint32_t d = __vadd2(dgA[r], s2);
      uint32_t h;
      if (SEMI) {
        // H-based opening without the serial H -> E -> H chain: with H = max(D, F, E),
        // E(t+1,q) = max(H - oe, E - e) = max(max(D, F) - oe, E - e) because oe >= e.  `e` enters a row as E(t,q) and
        // leaves it as E(t+1,q); the parked / handed-over E is the NEXT row's.
        st.fA[r] = __viaddmax_s16x2(st.glA[r], neg_oe2, st.fA[r]);    // F(t,q) = max(H(t,q-1) - oe, F(t,q-1) - e), biased frame
        h = __vimax3_s16x2(d, st.fA[r], __vadd2(eA, neg_oe2));        // eA holds E + open: the un-shift is an FMA-pipe add off the E chain
        eA = __vimax3_s16x2(eA, d, st.fA[r]);                         // E(t+1,q) + open = max(E + open, max(D, F)), biased frame
        st.glA[r] = h;
      } else {
        h = __vimax3_s16x2(d, st.fA[r], eA);
        st.fA[r] = __viaddmax_s16x2(d, neg_oe2, st.fA[r]);            // F(t,q+1) = max(F - e, D - oe), biased frame
        eA = __viaddmax_s16x2(d, neg_oe2, eA);                        // E(t+1,q) = max(E - e, D - oe), biased frame
      }
      hhA[r] = h; if (!SEMI) st.pA[r] = upA; upA = h;
    }
    {
      uint32_t s2 = prmt(rowC[r], rowD[r], st.selB);
      if (NMASK) s2 = mask_sel(s2, st.mB, nsc2);
      const uint32_t d = __vadd2(dgB[r], s2);
      uint32_t h;
      if (SEMI) {
        st.fB[r] = __viaddmax_s16x2(st.glB[r], neg_oe2, st.fB[r]);
        h = __vimax3_s16x2(d, st.fB[r], __vadd2(eB, neg_oe2));
        eB = __vimax3_s16x2(eB, d, st.fB[r]);
        st.glB[r] = h;
      } else {
        h = __vimax3_s16x2(d, st.fB[r], eB);
        st.fB[r] = __viaddmax_s16x2(d, neg_oe2, st.fB[r]);
        eB = __viaddmax_s16x2(d, neg_oe2, eB);
      }
      hhB[r] = h; if (!SEMI) st.pB[r] = upB; upB = h;
    }
  }
  if (BROW) {
    // first maximum of each lane's bottom row over the true query columns (lane i is on column col0 + k - i)
    uint32_t vA = __vsub2(upA, bt->biasA), vB = __vsub2(upB, bt->biasB);
    const uint32_t e2 = splat16(bias_ext), nq = (uint32_t)bt->nq;
    bt->biasA = __vadd2(bt->biasA, e2); bt->biasB = __vadd2(bt->biasB, e2);
    const uint32_t kA = ((uint32_t)k < nq ? 0x0000FFFFu : 0u) | ((uint32_t)(k - 1) < nq ? 0xFFFF0000u : 0u);
    const uint32_t kB = ((uint32_t)(k - 2) < nq ? 0x0000FFFFu : 0u) | ((uint32_t)(k - 3) < nq ? 0xFFFF0000u : 0u);
    vA = (vA & kA) | (0x80008000u & ~kA); vB = (vB & kB) | (0x80008000u & ~kB);
    track_update(bt->a, vA, vA, (uint32_t)k);
    track_update(bt->b, vB, vB, (uint32_t)k);
  }
  upA = __vadd2(upA, conv8); eA = __vadd2(eA, conv8); upB = __vadd2(upB, conv8); eB = __vadd2(eB, conv8);   // bottom row -> next strip's frame
  HE[(k - 2) * kFastThreads] = prmt(upB, eB, 0x7632);
  st.h8A = upA; st.eoA = eA; st.h8B = upB; st.eoB = eB;
  st.selB = st.selA_prev; st.selA_prev = selA;
  if (NMASK) { st.mB = st.mA_prev; st.mA_prev = mA; }
  if (RARE) {
    if (k <= 2) {
      // lanes that just ran a phantom column (A.H: step 0; B.L: 0,1; B.H: 0,1,2) get their strip-start state back
#pragma unroll
      for (int r = 0; r < 8; r++) {
        if (k == 0) { st.pA[r] = prmt(st.pA[r], pinitA[r], 0x7610); st.fA[r] = prmt(st.fA[r], finitA[r], 0x7610); if (SEMI) st.glA[r] = prmt(st.glA[r], hpinitA[r], 0x7610); }
        if (k <= 1) { st.pB[r] = pinitB[r]; st.fB[r] = finitB[r]; if (SEMI) st.glB[r] = hpinitB[r]; }
        else { st.pB[r] = prmt(st.pB[r], pinitB[r], 0x7610); st.fB[r] = prmt(st.fB[r], finitB[r], 0x7610); if (SEMI) st.glB[r] = prmt(st.glB[r], hpinitB[r], 0x7610); }
      }
    }
    const int li = k - kq;
    if (li >= 0 && li <= 3) {
      int hv[8];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const uint32_t w = (li & 2) ? hhB[r] : hhA[r];
        hv[r] = ((li & 1) ? lane_hi(w) : lane_lo(w)) - (col0 + kq + r) * bias_ext;     // back to the unbiased score (this lane is on column col0 + kq)
      }
      if (SEMI && !tr.capture) {
        // TAIL == TARGET: the last true query column, rows inside the target, strict '>' top-down (:160-178); the start
        // pass stops between strips once the forward score is reached (:301)
        if (!REV || li == 0 || tr.best < tr.stop) {
          const int rb = rowbase + 8 * li;
          if (rb < tlen) tr.last = li;
#pragma unroll
          for (int r = 0; r < 8; r++) if (rb + r < tlen && hv[r] > tr.best - tr.tie) { tr.best = hv[r]; tr.y = rb + r; }
        }
      } else if (li == caplane) {
#pragma unroll
        for (int r = 0; r < 8; r++) if (r == caprow) tr.cap = hv[r];
      }
    }
  }
}

// ---- quad-strip variant of nc_sweep (see local_sweep4): strips 4j..4j+3 as A.L, A.H, B.L, B.H, skewed by one column each.
// The rare per-step work (re-imposing the strip-start state of the late lanes after their phantom columns, steps 0..2; the scan of
// the last true query column / the global capture, steps kq..kq+3) runs in loops of its own, so the hot loop carries no rare code
// and can be unrolled (the diagonal / H rotations become register renaming).
//   kq      : step at which lane A.L is on the scanned / captured column (qlen - 1); lane i is there at step kq + i
template <int ALGO, bool REV, bool NMASK = false, bool BROW = false>
__device__ __forceinline__ void nc_sweep4(uint32_t* __restrict__ HE, const uint8_t* __restrict__ V, int nsteps,
                                          const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                          const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                          const uint32_t (&pinitA)[8], const uint32_t (&pinitB)[8],
                                          uint32_t neg_oe2, uint32_t neg_ext2,
                                          int kq, int rowbase, int tlen, int caplane, int caprow, NcTrack& tr, uint32_t nsc2 = 0,
                                          int col0 = 0, bool closed_left = false, const uint32_t* hpA = nullptr, const uint32_t* hpB = nullptr,
                                          const uint32_t* fpA = nullptr, const uint32_t* fpB = nullptr, NcBottom* bt = nullptr) {
  // hpA / hpB (nvb_strip.cuh GLOBAL; charged left column of the semi-global HEAD variants): H of the left boundary column per strip
  // row and lane, biased, instead of the free column; fpA / fpB: F of that column (default: -infinity)
  // col0 / closed_left (pruned start pass): HE / V point at column col0 of the matrix (kq is relative to it; the frame of the parked
  // values is absolute, so only the un-biasing of the scanned column needs col0), and what lies left of it reads as -infinity
  // instead of the free left column.
  constexpr bool SEMI = (ALGO == ALGO_SEMI_GLOBAL);
  const uint32_t neg2 = splat16(kNegSafe);
  // Both recurrences run in a BIASED frame: every H/E/F of (strip row r, column q) carries + (q + r) * ext.  Gap extension is
  // then free (semi-global: F' = max(F, Hleft - open), E' = max(E, M - open); global: F' = max(F, D - open), E' = max(E, D - open):
  // one VIADDMNMX each, no separate "- oe" adds), the diagonal step costs + 2 ext, which the caller folds into the substitution
  // tables, and a value that leaves a strip through its bottom row re-enters the next strip's frame with - 8 ext.  `neg_oe2` is
  // -gap_open here; for SEMI glA/glB hold H of the previous column (no add).  Left column of a semi-global strip: H(r, -1) = 0 ->
  // (r - 1) ext; the diagonal inputs (pinit) carry (r - 2) ext.  The -infinity floors stay unbiased: they only ever lose a max.
  const int bias_ext = -lane_lo(neg_ext2);
  const uint32_t conv8 = splat16(-8 * bias_ext);
  uint32_t hpinitA[8], hpinitB[8], finitA[8], finitB[8], pA0[8], pB0[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const uint32_t hp = SEMI ? splat16((r - 1) * bias_ext) : neg_oe2;
    hpinitA[r] = closed_left ? neg2 : (hpA ? hpA[r] : hp); hpinitB[r] = closed_left ? neg2 : (hpB ? hpB[r] : hp);
    finitA[r] = (fpA && !closed_left) ? fpA[r] : neg2; finitB[r] = (fpB && !closed_left) ? fpB[r] : neg2;
    pA0[r] = closed_left ? neg2 : pinitA[r]; pB0[r] = closed_left ? neg2 : pinitB[r];
  }
  NcState st;
#pragma unroll
  for (int r = 0; r < 8; r++) { st.pA[r] = pA0[r]; st.fA[r] = finitA[r]; st.glA[r] = hpinitA[r]; st.pB[r] = pB0[r]; st.fB[r] = finitB[r]; st.glB[r] = hpinitB[r]; }
  st.h8A = 0; st.eoA = 0; st.h8B = 0; st.eoB = 0;
  st.prevsh = 0xCC00u; st.selB = 0xCCCCu; st.selA_prev = 0xCCCCu;
  st.he = HE[kFastThreads]; st.v = V[kFastThreads];
  st.ncm_prev = 0; st.mB = 0; st.mA_prev = 0;
#define NC_STEP(RARE_, K_) nc_step<ALGO, REV, NMASK, RARE_, BROW>(st, K_, HE, V, rowA, rowB, rowC, rowD, pA0, pB0, hpinitA, hpinitB, finitA, finitB, neg_oe2, neg2, conv8, bias_ext, nsc2, \
                                                                 kq, rowbase, tlen, caplane, caprow, tr, col0, bt)
#if GB200_NC_SPLIT
  // segments: rare [0, r0e) -- hot [r0e, h0e) -- rare [h0e, r1e) -- hot [r1e, nsteps)
  const int r0e = min(3, nsteps);
  const int h0e = min(max(kq, r0e), nsteps);
  const int r1e = min(max(kq + 4, h0e), nsteps);
  int k = 0;
#pragma unroll 1
  for (; k < r0e; k++) NC_STEP(true, k);
#pragma unroll (SEMI ? kNcUnrollSemi : 2)
  for (; k < h0e; k++) NC_STEP(false, k);
#pragma unroll 1
  for (; k < r1e; k++) NC_STEP(true, k);
#pragma unroll 1
  for (; k < nsteps; k++) NC_STEP(false, k);
#else
#pragma unroll (SEMI ? 1 : 2)
  for (int k = 0; k < nsteps; k++) NC_STEP(true, k);
#endif
#undef NC_STEP
}

// selector bytes + top boundary of one pass into shared memory.  mode 0: forward words; mode 2: true-length reversed.
// SEMI: the parked E is the first row's own E(0,c) = max(H(-1,c) - oe, E(-1,c) - e) (see nc_sweep); otherwise E(-1,c).
// floor_from: columns >= floor_from get the -infinity floor instead of the top boundary (pruned start pass: the first quad-strip never
// reads them, and for the later ones the cell above such a column is outside every path that reaches the forward score)
template <bool TOP_ZERO, bool SEMI = true, bool NMASK = false>
__device__ __forceinline__ bool nc_prologue(uint32_t* HE, uint8_t* V, const uint32_t* qw, uint32_t qlen, int QR, int mode, int o, int ext, uint32_t nval = 0,
                                            int floor_from = 0x7FFFFFFF) {
  bool odd = false;
  const int ncol = QR * 8;
  V[0] = 0x88; HE[0] = pack2(0, kNegSafe);
  for (int j = 0; j < QR; j++) {
    const uint32_t w = (mode == 2) ? rev_true_word(qw, (int)qlen, j) : qw[j];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int c = 8 * j + k;
      V[(c + 1) * kFastThreads] = col_selector<NMASK>(nib(w, k), (uint32_t)c >= qlen, nval, odd);
      // top boundary: H(-1, c), E = -inf   (semiglobal_kernel_template.h:87-99, global.h:57-60)
      if (SEMI) {
        const int htop = TOP_ZERO ? 0 : (c == 0 ? 0 : -(o + ext * c));                  // free top row: (H,E)(-1,c) = (0,0)
        // biased frame of nc_sweep4: H(-1,c) carries (c - 1) ext, E(0,c) carries c ext
        // ... and E is kept as E + open (nc_sweep4 takes max(E + open, D, F) in one VIMNMX3)
        HE[(c + 1) * kFastThreads] = (c >= floor_from) ? pack2(kNegSafe, kNegSafe)
                                                       : pack2(htop + (c - 1) * ext, (TOP_ZERO ? -ext : htop - (o + ext)) + c * ext + o);
      } else {
        HE[(c + 1) * kFastThreads] = pack2((c == 0 ? 0 : -(o + ext * c)) + (c - 1) * ext, kNegSafe);   // H(-1,c) in the biased frame
      }
    }
  }
#pragma unroll
  for (int x = 1; x <= 4; x++) { V[(ncol + x) * kFastThreads] = 0x88; HE[(ncol + x) * kFastThreads] = pack2(0, kNegSafe); }
  return odd;
}

__device__ __forceinline__ bool target_has_odd_base(const uint32_t* tw, uint32_t tlen, int TR) {
  return target_odd_scan(tw, tlen, TR);
}

// TAILQ: TAIL in {QUERY, BOTH} -- the bottom row of the last processed (padded) strip is scanned as well (NcBottom); its own
// instantiations, so that the default kernels carry none of it.
template <int START, bool NMASK = false, bool TAILQ = false>
__global__ void GB200_FAST_LB k_semiglobal_s16(const FastArgs fa) {
  extern __shared__ uint32_t smem[];
  const DpArgs& a = fa.a;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (NMASK) { if (tid >= fa.work_count[0]) return; tid = fa.worklist[tid]; if (fa.flags[tid] & 2u) return; }   // N-aware pass: one thread per listed pair
  else { if (tid >= (uint32_t)a.n) return; if (fa.flags && fa.flags[tid]) return; }   // flagged pairs belong to the other passes (k_classify)
  uint32_t* HE = smem + 2 * kFastThreads + threadIdx.x;          // slot s holds column s - 1; slots -2..0 are dummies
  uint8_t* V = (uint8_t*)(smem + (fa.q8max + kSlotPad) * kFastThreads) + threadIdx.x;
  const uint32_t* qw = a.pq + (a.q_offs[tid] >> 3);
  const uint32_t* tw = a.pt + (a.t_offs[tid] >> 3);
  const uint32_t qlen = a.q_lens[tid], tlen = a.t_lens[tid];
  const int QR = (int)n_words(qlen), TR = (int)n_words(tlen);
  // biased frame (nc_sweep4): a diagonal step moves one row and one column, so both substitution scores carry + 2 ext
  const uint32_t mis4 = ((uint32_t)(-a.mismatch + 2 * a.gap_extend) & 0xFFu) * 0x01010101u;
  const uint32_t delta = ((uint32_t)(a.match + 2 * a.gap_extend) ^ (uint32_t)(-a.mismatch + 2 * a.gap_extend)) & 0xFFu;
  const uint32_t neg_oe2 = splat16(-a.gap_open), neg_ext2 = splat16(-a.gap_extend);      // gap OPEN only: extension is free in the biased frame
  const uint32_t nval = a.n_value, nsc2 = splat16(a.n_score + 2 * a.gap_extend), nsc4 = ((uint32_t)(a.n_score + 2 * a.gap_extend) & 0xFFu) * 0x01010101u;
  const bool top_zero = (a.head == SKIP_BOTH || a.head == SKIP_QUERY);          // HEAD in {QUERY,BOTH}: free top row (:87)
  const bool head_t = (a.head == SKIP_TARGET || a.head == SKIP_BOTH);           // HEAD in {TARGET,BOTH}: free left column (:114-121)
  const int ncol = QR * 8, nsteps = ncol + 2 + (int)fa.one;

  // START_REV: the start pass as a launch of its own behind the score-only kernel (the two passes as ONE kernel keep each other's
  // state live: the forward sweep of the fused kernel runs 4 - 10 % slower than the same sweep alone; measured)
  constexpr bool FWD = (START != START_REV), REVP = (START == START_POS || START == START_REV);
  if (FWD) {
    bool odd = top_zero ? nc_prologue<true, true, NMASK>(HE, V, qw, qlen, QR, 0, a.gap_open, a.gap_extend, nval)
                        : nc_prologue<false, true, NMASK>(HE, V, qw, qlen, QR, 0, a.gap_open, a.gap_extend, nval);
    if (odd || qlen == 0) {   // cannot happen (see k_classify); safety net
      const uint32_t slot = atomicAdd(fa.work_count + 1, 1u);
      fa.worklist2[slot] = tid;
      return;
    }
  } else if (qlen == 0) return;
  uint32_t pinit[8];
#pragma unroll
  for (int r = 0; r < 8; r++) pinit[r] = splat16((r - 2) * a.gap_extend);   // HEAD in {TARGET,BOTH}: free left column, diag H(r-1,-1) = 0 (:114-121), biased
  // HEAD in {NONE,QUERY}: the left column is charged (:123-128): H(t, -1) = -(o + e t), and the diagonal input of row t in column 0 is
  // the SAME value (0 for t = 0), not H(t - 1, -1) -- t counted from the first strip of the PASS (the start pass restarts the count).
  // The sweep takes the diagonal input of strip row r from the left neighbour of row r - 1, so the left-neighbour column it is given
  // is the DIAGONAL column shifted up by one row (entry r = the diagonal input of row r + 1 = -(o + e (t + 1))), row 0's diagonal
  // comes through pinit, and what the true left neighbour contributes to column 0 -- F(t, 0) = H(t, -1) - oe -- is handed over as
  // the F of the boundary column.  Biased frame: the diagonal input sits in the frame of (row r - 1, column -1), left neighbour and
  // F in (row r, column -1); F(t, 0) + r e = [H(t, -1) + (r - 1) e] - open.
  auto left_column = [&](int t0, uint32_t (&pA)[8], uint32_t (&pB)[8], uint32_t (&hA)[8], uint32_t (&hB)[8], uint32_t (&fA)[8], uint32_t (&fB)[8]) {
    const int o = a.gap_open, ext = a.gap_extend;
    if (head_t) {                       // free left column: H(t, -1) = 0, F = -infinity (one sweep call site serves both cases)
#pragma unroll
      for (int r = 0; r < 8; r++) { pA[r] = pB[r] = pinit[r]; hA[r] = hB[r] = splat16((r - 1) * ext); fA[r] = fB[r] = splat16(kNegSafe); }
      return;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int t = t0 + r, fd = (r - 2) * ext, fh = (r - 1) * ext;
      pA[r] = pack2((t == 0 ? 0 : -(o + ext * t)) + fd, -(o + ext * (t + 8)) + fd);
      pB[r] = pack2(-(o + ext * (t + 16)) + fd, -(o + ext * (t + 24)) + fd);
      hA[r] = pack2(-(o + ext * (t + 1)) + fh, -(o + ext * (t + 9)) + fh);
      hB[r] = pack2(-(o + ext * (t + 17)) + fh, -(o + ext * (t + 25)) + fh);
      fA[r] = pack2(-(o + ext * t) + fh - o, -(o + ext * (t + 8)) + fh - o);
      fB[r] = pack2(-(o + ext * (t + 16)) + fh - o, -(o + ext * (t + 24)) + fh - o);
    }
  };

  NcTrack tr; tr.best = kMinusInf; tr.y = (int)qlen; tr.stop = 0; tr.cap = 0;       // :63-64
  NcBottom bot;
  const bool tail_t = !TAILQ || a.tail == SKIP_BOTH;                                // the last query column is scanned (:160-178)
  const int kq_scan = tail_t ? (int)qlen - 1 : -1000;
  const int NQ = (TR + 3) >> 2;
  for (int j = 0; FWD && j < NQ; j++) {
    uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
    const uint32_t wa = tw[4 * j];
    const uint32_t wb = (4 * j + 1 < TR) ? tw[4 * j + 1] : 0u;
    const uint32_t wc = (4 * j + 2 < TR) ? tw[4 * j + 2] : 0u;
    const uint32_t wd = (4 * j + 3 < TR) ? tw[4 * j + 3] : 0u;
    quad_tables<NMASK || TAILQ>(wa, wb, wc, wd, mis4, delta, nval, nsc4, rowA, rowB, rowC, rowD);   // TAILQ: the pad rows of the last strip count, and the pads are N
    // Measured (C2, 262144 pairs): handing the free column's start state to the sweep as run-time arrays (one call site for both HEAD
    // kinds) makes the score-only kernel faster (3.87 -> 3.71 ms) and the kernel with the start pass slower (4.59 -> 4.91 ms: the arrays
    // stay live next to the second pass's state), so the latter keeps a call site of its own for the free column.
    if (START == START_POS && !TAILQ && head_t) {      // (fused kernel only)
      nc_sweep4<ALGO_SEMI_GLOBAL, false, NMASK>(HE, V, nsteps, rowA, rowB, rowC, rowD, pinit, pinit, neg_oe2, neg_ext2, kq_scan, 32 * j, (int)tlen, 0, 0, tr, nsc2);
      continue;
    }
    uint32_t pA[8], pB[8], hA[8], hB[8], fA[8], fB[8];
    left_column(32 * j, pA, pB, hA, hB, fA, fB);
    if (TAILQ && j == NQ - 1) {
      // last quad-strip: keep the first maximum of every lane's bottom row (the lane of strip TR - 1 is read below)
      ncbottom_init(bot, 0, (int)qlen, a.gap_extend);
      nc_sweep4<ALGO_SEMI_GLOBAL, false, NMASK, true>(HE, V, nsteps, rowA, rowB, rowC, rowD, pA, pB, neg_oe2, neg_ext2, kq_scan, 32 * j, (int)tlen, 0, 0, tr, nsc2,
                                                      0, false, hA, hB, fA, fB, &bot);
    } else nc_sweep4<ALGO_SEMI_GLOBAL, false, NMASK>(HE, V, nsteps, rowA, rowB, rowC, rowD, pA, pB, neg_oe2, neg_ext2, kq_scan, 32 * j, (int)tlen, 0, 0, tr, nsc2,
                                                     0, false, hA, hB, fA, fB);
  }
  int tr_x = (int)tlen;                                                             // query_batch_end = ref_len unless the bottom-row scan finds a column (Q8)
  if (FWD) {
    if (TAILQ) {                                                                    // :185-216, strict '>' after the column scan; :214 `if (maxXY_x != ref_len)`
      int v, m; ncbottom_get(bot, (TR - 1) & 3, 0, v, m);
      if (v > tr.best) { tr.best = v; tr_x = m; }
      if (tr_x != (int)tlen) tr.y = (int)qlen;
    }
    a.score[tid] = tr.best; a.t_end[tid] = tr.y; a.q_end[tid] = tr_x;
  } else { tr.best = a.score[tid]; tr.y = a.t_end[tid]; tr_x = a.q_end[tid]; }      // the forward launch's results

  if (REVP) {
    // start pass (:227-384): the same DP over the TRUE-LENGTH reversed sequences, from the strip that holds the
    // end row, until the running best reaches the forward score
    const int fwd = tr.best;
    const uint32_t ua = (uint32_t)TR - (uint32_t)((tr.y >> 3) + 1);
    const uint32_t gend_reg = ua > 0 ? ua - 1 : ua;                                   // :273 (unsigned in the reference)
    // Score-bound pruning (exact; see the LOCAL start pass): a path through the cell (row i of this pass, column c) with d = c - i > 0
    // has at most i + 1 diagonal steps before that cell and qlen - 1 - c after it, and at least d of its columns sit in gaps, which
    // cost at least open + (d - 1) ext (the cheapest case is one gap on the top boundary, quirk Q4), so it scores at most
    // m (qlen - d) - open - (d - 1) ext, m = the best substitution score.  No path that reaches the forward score enters the columns
    // c > i + slack, slack = (m qlen - score - open + ext) / (m + ext): quad-strip k of the pass stops after column 32 (k + 1) + slack,
    // the cells beyond act as -infinity (a lower bound of their true value; every cell on a path that reaches the score keeps exact
    // inputs, no cell exceeds its true value), and rows whose last column was not computed cannot hold the score.  If the pruned
    // pass ends WITHOUT reaching the forward score (possible when the forward path starts with a gap on the top boundary, which the
    // reversed DP prices differently), it is run again unpruned, so that the reference's "first best row" is reproduced as well.
    // HEAD == BOTH has a free top row: no bound, no pruning.  (slack is warp-uniform, the maximum over the converged lanes.)
    const int mstep = NMASK ? max(a.match, a.n_score) : a.match;
    int slack = ncol;
    if (GB200_PRUNE_REVERSE && !TAILQ && !top_zero && head_t && mstep + a.gap_extend > 0) {       // (a charged left column / a bottom-row scan: no such bound either)
      const int num = mstep * (int)qlen - fwd - a.gap_open + a.gap_extend;
      slack = (int)__reduce_max_sync(__activemask(), (unsigned)max(num, 0) / (unsigned)(mstep + a.gap_extend));
      // The bound pass cannot reach the forward score when the forward path starts with a gap on the top boundary (the reversed DP
      // prices that gap one extension higher), and then the exact pass has to follow: twice the work.  Such a path starts on target
      // row 0, i.e. it spends tr.y + 1 - qlen deletions or more, which the score allows only up to (m qlen - score - open) / ext.
      // Where several pairs of a warp are in that position (reads in tight windows: the reference's real-read fixture), the warp
      // runs the exact pass straight away; a few such pairs (C2: reads anywhere in a 600 bp window) are not worth giving the bound up.
      const int maxdel = a.gap_extend > 0 ? max(mstep * (int)qlen - fwd - a.gap_open, 0) / a.gap_extend : 0x7FFFFFF;
      const bool risky = (tr.y + 1 - (int)qlen) <= maxdel;
      if (__popc(__ballot_sync(__activemask(), risky)) >= GB200_PRUNE_RISKY_LANES) slack = ncol;
    }
    // The other side of the band.  A path of this pass that reaches the forward score is the mirror of a forward alignment that ends
    // on a row t' with H(t', qlen - 1) = score; the forward scan recorded the FIRST such row (strict '>' top-down) and this pass starts
    // at most one strip before the strip that holds it, so such a path leaves the left boundary within the first 16 rows of the pass
    // (or comes off the top boundary, which is tighter still).  Through the cell (row i, column c) with i - c = d > 0 it has spent at
    // least d - 15 rows in vertical gaps: score <= m qlen - open - (d - 15) ext  ->  d <= 15 + (m qlen - score - open) / ext (15 when
    // not even one gap is affordable; no bound when extension is free).  Quad-strip k sweeps the columns [32 k - slackL, 32 (k + 1) +
    // slack); what lies left of the window reads as -infinity (nc_sweep4 closed_left), not as the free left column.
    int slackL = 1 << 24;
    if (GB200_PRUNE_REVERSE && GB200_PRUNE_REVERSE_LOW && slack < ncol) {
      const int numL = mstep * (int)qlen - fwd - a.gap_open;
      const int lo = numL < 0 ? 15 : (a.gap_extend > 0 ? 15 + numL / a.gap_extend : (1 << 24));
      slackL = (int)__reduce_max_sync(__activemask(), (unsigned)lo);
    }
    const int c_max = (int)((qlen - 1) & ~3u);                                          // the scanned column stays inside every window
    NcTrack rt;
    int bot_v = 0, bot_m = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
      const int lim0 = min(ncol, (32 + slack + 3) & ~3);
      if (top_zero) nc_prologue<true, true, NMASK>(HE, V, qw, qlen, QR, 2, a.gap_open, a.gap_extend, nval);
      else nc_prologue<false, true, NMASK>(HE, V, qw, qlen, QR, 2, a.gap_open, a.gap_extend, nval, lim0 < ncol ? lim0 : 0x7FFFFFFF);
      rt.best = kMinusInf; rt.y = 0; rt.stop = fwd; rt.cap = 0;                         // :275-276
      int lim = lim0, row_lo = 0;
      // TAILQ: bottom row of the last strip the reference processes; no strip at all: the top boundary, whose first maximum is column 0, value 0
      if (TAILQ) { bot_v = 0; bot_m = 0; }
      for (uint32_t i = gend_reg; i < (uint32_t)TR && rt.best < fwd; i += 4, row_lo += 32) {
        uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
        const uint32_t wa = rev_true_word(tw, (int)tlen, (int)i);
        const uint32_t wb = (i + 1 < (uint32_t)TR) ? rev_true_word(tw, (int)tlen, (int)i + 1) : 0u;
        const uint32_t wc = (i + 2 < (uint32_t)TR) ? rev_true_word(tw, (int)tlen, (int)i + 2) : 0u;
        const uint32_t wd = (i + 3 < (uint32_t)TR) ? rev_true_word(tw, (int)tlen, (int)i + 3) : 0u;
        quad_tables<NMASK || TAILQ>(wa, wb, wc, wd, mis4, delta, nval, nsc4, rowA, rowB, rowC, rowD);   // TAILQ: the pad rows of the last strip count, and the pads are N
        // a truncated sweep computes the columns < lim in all four lanes; if it stops on the scanned column (qlen - 1 >= lim) the lanes
        // that get there only see values below the forward score, which never decide the result
        const int c_lo = (slack < ncol) ? min(max(row_lo - slackL, 0) & ~3, c_max) : 0;
        const int ns = ((lim < ncol) ? lim + 2 + (int)fa.one : nsteps) - c_lo;
        if (!TAILQ && head_t) {                       // (short pruned sweeps: the constant start state of the free column is cheaper to set up)
          nc_sweep4<ALGO_SEMI_GLOBAL, true, NMASK>(HE + (size_t)c_lo * kFastThreads, V + (size_t)c_lo * kFastThreads, ns, rowA, rowB, rowC, rowD, pinit, pinit,
                                                   neg_oe2, neg_ext2, (int)qlen - 1 - c_lo, 8 * (int)i, (int)tlen, 0, 0, rt, nsc2, c_lo, c_lo > 0);
          lim = min(ncol, lim + 32);
          continue;
        }
        uint32_t pA[8], pB[8], hA[8], hB[8], fA[8], fB[8];
        left_column(row_lo, pA, pB, hA, hB, fA, fB);                                    // rows counted from the first strip of this pass
        if (TAILQ) {
          ncbottom_init(bot, 0, (int)qlen, a.gap_extend);
          rt.last = 0;
          nc_sweep4<ALGO_SEMI_GLOBAL, true, NMASK, true>(HE, V, nsteps, rowA, rowB, rowC, rowD, pA, pB, neg_oe2, neg_ext2, kq_scan, 8 * (int)i, (int)tlen, 0, 0, rt, nsc2,
                                                         0, false, hA, hB, fA, fB, &bot);
          // the last strip of this quad-strip that the reference processes: with the column scan, the last lane whose scan ran (the pass
          // stops between strips once the forward score is reached); without it, every strip up to TR - 1
          const int last = tail_t ? rt.last : min(3, TR - 1 - (int)i);
          ncbottom_get(bot, last, 0, bot_v, bot_m);
        } else nc_sweep4<ALGO_SEMI_GLOBAL, true, NMASK>(HE, V, nsteps, rowA, rowB, rowC, rowD, pA, pB, neg_oe2, neg_ext2, (int)qlen - 1, 8 * (int)i, (int)tlen, 0, 0, rt, nsc2,
                                                        0, false, hA, hB, fA, fB);
        lim = min(ncol, lim + 32);
      }
      if (rt.best >= fwd || slack >= ncol) break;
      slack = ncol;                                                                    // not reached under the bound: the exact, unpruned pass
    }
    int rt_x = tr_x;                                                                  // (maxXY_x is not reset between the passes)
    if (TAILQ) {
      if (bot_v > rt.best) { rt.best = bot_v; rt_x = bot_m; }
      if (rt_x != (int)tlen) rt.y = (int)qlen;
    }
    a.t_start[tid] = ((int)tlen - 1) - rt.y;                                          // :380-381
    a.q_start[tid] = ((int)qlen - 1) - rt_x;
  }
}

// SEMI_GLOBAL with TAIL = NONE: neither the last query column nor the last target row is ever scanned (semiglobal_kernel_template.h:
// the scans of :160-178 and :185-216 are the only writers of maxHH), so every pair reports the initial state of the reference's
// trackers: score MINUS_INF, target end = read_len, query end = ref_len (:63-64); the start pass stops before its first strip
// (:301: maxHH >= the "forward score") and reports ref_len - 1 and read_len - 1 - ref_len (:380-381).  No DP runs.
template <int START>
__global__ void __launch_bounds__(256) k_semiglobal_tail_none(const DpArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (uint32_t)a.n) return;
  const int qlen = (int)a.q_lens[tid], tlen = (int)a.t_lens[tid];
  a.score[tid] = kMinusInf; a.t_end[tid] = qlen; a.q_end[tid] = tlen;
  if (START == START_POS) { a.t_start[tid] = tlen - 1; a.q_start[tid] = qlen - 1 - tlen; }
}

__global__ void GB200_FAST_LB k_global_s16(const FastArgs fa) {
  extern __shared__ uint32_t smem[];
  const DpArgs& a = fa.a;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (uint32_t)a.n) return;
  if (fa.flags && fa.flags[tid]) return;                         // general kernel (k_classify)
  uint32_t* HE = smem + 2 * kFastThreads + threadIdx.x;          // slot s holds column s - 1; slots -2..0 are dummies
  uint8_t* V = (uint8_t*)(smem + (fa.q8max + kSlotPad) * kFastThreads) + threadIdx.x;
  const uint32_t* qw = a.pq + (a.q_offs[tid] >> 3);
  const uint32_t* tw = a.pt + (a.t_offs[tid] >> 3);
  const uint32_t qlen = a.q_lens[tid], tlen = a.t_lens[tid];
  const int QR = (int)n_words(qlen), TR = (int)n_words(tlen);
  const int o = a.gap_open, ext = a.gap_extend;
  // biased frame (nc_sweep4): both substitution scores carry + 2 ext, gaps cost their OPEN part only
  const uint32_t mis4 = ((uint32_t)(-a.mismatch + 2 * ext) & 0xFFu) * 0x01010101u;
  const uint32_t delta = ((uint32_t)(a.match + 2 * ext) ^ (uint32_t)(-a.mismatch + 2 * ext)) & 0xFFu;
  const uint32_t neg_oe2 = splat16(-o), neg_ext2 = splat16(-ext);
  const int ncol = QR * 8, nsteps = ncol + 2 + (int)fa.one;
  bool odd = nc_prologue<false, false>(HE, V, qw, qlen, QR, 0, o, ext);
  if (odd || qlen == 0 || tlen == 0) {   // cannot happen (see k_classify); safety net
    const uint32_t slot = atomicAdd(fa.work_count + 1, 1u);
    fa.worklist2[slot] = tid;
    return;
  }
  NcTrack tr; tr.best = 0; tr.y = 0; tr.stop = 0; tr.cap = 0;
  const int NQ = (TR + 3) >> 2;
  const int last_strip = (int)(tlen - 1) >> 3, caprow = (int)(tlen - 1) & 7;
  for (int j = 0; j < NQ; j++) {
    uint32_t rowA[8], rowB[8], rowC[8], rowD[8], pinitA[8], pinitB[8];
    const uint32_t wa = tw[4 * j];
    const uint32_t wb = (4 * j + 1 < TR) ? tw[4 * j + 1] : 0u;
    const uint32_t wc = (4 * j + 2 < TR) ? tw[4 * j + 2] : 0u;
    const uint32_t wd = (4 * j + 3 < TR) ? tw[4 * j + 3] : 0u;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      rowA[r] = row_table(base_idx(nib(wa, r)), mis4, delta); rowB[r] = row_table(base_idx(nib(wb, r)), mis4, delta);
      rowC[r] = row_table(base_idx(nib(wc, r)), mis4, delta); rowD[r] = row_table(base_idx(nib(wd, r)), mis4, delta);
      const int t0 = 32 * j + r;                                                     // diag(t) = t ? -(o + e*t) : 0   (global.h:66-70)
      const int fb = (r - 2) * ext;                                                   // frame of (row r - 1, column -1)
      pinitA[r] = pack2((t0 == 0 ? 0 : -(o + ext * t0)) + fb, -(o + ext * (t0 + 8)) + fb);
      pinitB[r] = pack2(-(o + ext * (t0 + 16)) + fb, -(o + ext * (t0 + 24)) + fb);
    }
    const bool has_last = (last_strip >> 2) == j;
    nc_sweep4<ALGO_GLOBAL, false>(HE, V, nsteps, rowA, rowB, rowC, rowD, pinitA, pinitB, neg_oe2, neg_ext2,
                                  has_last ? (int)qlen - 1 : -1000, 0, (int)tlen, last_strip & 3, caprow, tr);
  }
  a.score[tid] = tr.cap;                                                              // H(ref_len-1, read_len-1)  (global.h:285-299)
}

// ---- general kernel over the fallback worklist ----------------------------------------------------
template <int ALGO, int START>
__global__ void __launch_bounds__(128) k_worklist(const FastArgs fa, int which) {
  const DpArgs& a = fa.a;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const uint32_t count = fa.work_count[which];
  const uint32_t* __restrict__ list = which ? fa.worklist2 : fa.worklist;
  const Sc sc = make_sc(a, ALGO == ALGO_GLOBAL);
  for (uint32_t w = gtid; w < count; w += nthreads) {
    const uint32_t pi = list[w];
    PairCtx p;
    p.qw = a.pq + (a.q_offs[pi] >> 3); p.tw = a.pt + (a.t_offs[pi] >> 3);
    p.qlen = a.q_lens[pi]; p.tlen = a.t_lens[pi]; p.QR = n_words(p.qlen); p.TR = n_words(p.tlen);
    p.cs = a.colstate + gtid; p.stride = a.col_stride;
    if (ALGO == ALGO_LOCAL) {
      int best, bq, bt;
      if (START == START_TB) { if (a.tb_banded) banded_tb_forward(a, sc, p, pi, best, bq, bt); else local_forward<true, false>(a, sc, p, pi, best, bq, bt); }
      else local_forward<false, false>(a, sc, p, pi, best, bq, bt);
      a.score[pi] = best; a.q_end[pi] = bq; a.t_end[pi] = bt;
      if (START == START_POS) { int sx, sy; local_reverse(sc, p, best, bq, bt, sx, sy); a.q_start[pi] = sx; a.t_start[pi] = sy; }
    } else if (ALGO == ALGO_BANDED) {
      int best, bq, bt;
      banded_forward(a, sc, p, best, bq, bt);
      a.score[pi] = best; a.q_end[pi] = bq; a.t_end[pi] = bt;
    } else if (ALGO == ALGO_SEMI_GLOBAL) {
      semiglobal_pair<START, false>(a, sc, p, pi);
    } else {
      a.score[pi] = global_forward<false>(a, sc, p, pi);
    }
  }
}

__global__ void k_zero_u32(uint32_t* p) { p[0] = 0; p[1] = 0; }

// ---- host side --------------------------------------------------------------------------------------
#ifndef GB200_BANDED_TB_S16
#define GB200_BANDED_TB_S16 1
#endif
#ifndef GB200_BANDED_S16
#define GB200_BANDED_S16 1
#endif
bool fast_path_eligible(const DpArgs& a, int algo, int start, bool second, uint32_t max_q, uint32_t max_t) {
  if (algo == ALGO_BANDED) {                     // banded traceback = LOCAL recurrence over the tile band; score-only BANDED = H-based opening (banded_sweep4)
    if (a.k_band < 8) return false;
    if (start == START_TB ? !GB200_BANDED_TB_S16 : !GB200_BANDED_S16) return false;
    algo = ALGO_LOCAL;                           // same value range: clamped at 0, keys score * 8 + 7
    if (start != START_TB) start = START_NONE;   // the banded kernel reports score + end cell whatever start_pos says (gasal_align.h:78-81)
  }
  if (second || (start == START_TB && algo != ALGO_LOCAL)) return false;
  const uint32_t q8 = (max_q + 7u) & ~7u, t8 = (max_t + 7u) & ~7u;
  if (q8 == 0 || q8 > kFastMaxQ8 || t8 == 0) return false;
  if (a.match < 0 || a.match > 127 || a.mismatch < 0 || a.mismatch > 127) return false;
  if (a.gap_open < 0 || a.gap_extend < 0 || a.gap_open + a.gap_extend > 2000) return false;
  const int64_t oe = (int64_t)a.gap_open + a.gap_extend;
  if (algo == ALGO_LOCAL) {
    // keys are score*8+7 in a signed 16-bit lane; E/F stay above -(mismatch + gap_open + gap_extend)
    return (uint64_t)std::min(q8, t8) * (uint64_t)a.match <= 4000;
  }
  if (algo == ALGO_SEMI_GLOBAL) {
    // every HEAD / TAIL variant: TAIL = TARGET scans the last query column, TAIL in {QUERY, BOTH} also / only the bottom row of the last
    // processed strip (k_semiglobal_s16<.., TAILQ>), TAIL = NONE scans nothing (the reference's result is its initial state, k_semiglobal_tail_none)
    if (a.tail == SKIP_NONE) return true;
    // the bottom-row scan reads PAD rows, and the pads of a forward sequence are N: their score must fit the byte tables
    if (a.tail != SKIP_TARGET && (a.n_score + 2 * a.gap_extend < -127 || a.n_score + 2 * a.gap_extend > 127)) return false;
    const bool head_t = (a.head == SKIP_TARGET || a.head == SKIP_BOTH);
    // every H/E/F is >= -(2*oe + ext*(q8+2) + mismatch) and <= q8*match: keep both well inside (kNegSafe, 32767); a charged left
    // column (HEAD in {NONE, QUERY}) adds ext * t8
    if (2 * oe + (int64_t)a.gap_extend * (q8 + 2 + (head_t ? 0 : t8 + 32)) + a.mismatch + 64 > 12000) return false;
    // the biased frame adds up to (q8 + 16) * ext to every value, and the substitution bytes carry + 2 ext
    if (a.match + 2 * a.gap_extend > 127) return false;
    return (int64_t)q8 * a.match + (int64_t)(q8 + 24) * a.gap_extend <= 30000;
  }
  if (algo == ALGO_GLOBAL) {
    if (start != START_NONE) return false;
    if (2 * oe + (int64_t)a.gap_extend * (q8 + t8 + 2) + a.mismatch + 64 > 12000) return false;
    if (a.match + 2 * a.gap_extend > 127) return false;                      // biased frame, as for semi-global
    return (int64_t)std::min(q8, t8) * a.match + (int64_t)(q8 + 24) * a.gap_extend <= 30000;
  }
  return false;
}

#ifndef GB200_LOCAL_G4
#define GB200_LOCAL_G4 1
#endif
// (the same split for LOCAL WITH_START was measured: 4.97 ms against 4.95 ms fused on C2 -- no gain, not kept)
#ifndef GB200_SEMI_SPLIT_START
#define GB200_SEMI_SPLIT_START 1     // semi-global WITH_START: score-only kernel + start-pass kernel (START_REV) instead of the fused kernel
#endif
int fast_path_launch(const DpArgs& a, int algo_in, int start, FastScratch& fs, uint32_t max_q, uint32_t max_t, bool maybe_odd, int force_passes, cudaStream_t str) {
  int algo = algo_in;
  if (algo == ALGO_SEMI_GLOBAL && a.tail == SKIP_NONE) {
    const uint32_t g = ((uint32_t)a.n + 255u) / 256u;
    if (start == START_POS) k_semiglobal_tail_none<START_POS><<<g, 256, 0, str>>>(a); else k_semiglobal_tail_none<START_NONE><<<g, 256, 0, str>>>(a);
    return 1;
  }
  const bool passes = maybe_odd || force_passes;        // false: nothing but A/C/G/T in this batch (host packer): one launch, no lists
  const uint32_t q8 = (max_q + 7u) & ~7u;
  if (fs.cap < (size_t)a.n) {
    if (fs.worklist) cudaFree(fs.worklist);
    if (fs.flags) cudaFree(fs.flags);
    fs.worklist = nullptr; fs.flags = nullptr;
    if (cudaMalloc(&fs.worklist, 2 * sizeof(uint32_t) * (size_t)a.n) != cudaSuccess) return -1;     // worklist, then worklist2
    if (cudaMalloc(&fs.flags, sizeof(uint32_t) * (size_t)a.n) != cudaSuccess) return -1;
    fs.cap = (size_t)a.n;
  }
  if (!fs.side) {
    if (cudaStreamCreateWithFlags(&fs.side, cudaStreamNonBlocking) != cudaSuccess) return -1;
    if (cudaEventCreateWithFlags(&fs.ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&fs.ev_join, cudaEventDisableTiming) != cudaSuccess) return -1;
  }
  if (!fs.work_count && cudaMalloc(&fs.work_count, 2 * sizeof(uint32_t)) != cudaSuccess) return -1;
  // Pairs with a base outside A/C/G/T: LOCAL / SEMI_GLOBAL re-run them on the N-aware variant of the same s16x2 kernel (one more
  // LOP3 per cell pair); what is left (codes other than N; GLOBAL, whose rule makes N an ordinary symbol) goes to the general
  // int32 kernel, grid-striding over the list with a grid sized for the machine, not for a fixed 64 CTAs.
  const int nbias = (algo == ALGO_SEMI_GLOBAL) ? 2 * a.gap_extend : 0;
  const bool banded = (algo == ALGO_BANDED);
  const bool banded_sc = banded && start != START_TB;     // score-only BANDED: H-based recurrence, no direction matrix
  if (banded) { algo = ALGO_LOCAL; if (banded_sc) start = START_NONE; }      // (fast_path_eligible): clamped recurrences, banded sweep
  const bool nmask = (algo == ALGO_LOCAL || algo == ALGO_SEMI_GLOBAL) && a.n_score + nbias >= -127 && a.n_score + nbias <= 127;
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint32_t n128 = ((uint32_t)a.n + 127u) & ~127u;
  const uint32_t slots = nmask ? std::min<uint32_t>(n128, kFastFallbackSlots) : std::min<uint32_t>(n128, (uint32_t)sms * 4u * 128u);
  const size_t cs_need = (size_t)slots * q8;
  if (fs.colstate_words < cs_need) {
    if (fs.colstate) cudaFree(fs.colstate);
    fs.colstate = nullptr;
    if (cudaMalloc(&fs.colstate, cs_need * sizeof(uint32_t)) != cudaSuccess) return -1;
    fs.colstate_words = cs_need;
  }
  FastArgs fa; fa.a = a; fa.q8max = q8; fa.worklist = fs.worklist; fa.worklist2 = fs.worklist + fs.cap; fa.work_count = fs.work_count; fa.flags = passes ? fs.flags : nullptr;
  fa.eight = 8; fa.thirtytwo = 32; fa.one = 1; fa.mm = 0x00204081u; fa.sixteen = 16;
  // GASALX_FAST_SMEM_PAD_KB (experiments only): extra dynamic shared memory per CTA, to measure the same kernel at a lower residency
  static const size_t smem_pad = [] { const char* e = getenv("GASALX_FAST_SMEM_PAD_KB"); return e ? (size_t)atoi(e) * 1024 : (size_t)0; }();
  const size_t smem = (size_t)(q8 + kSlotPad) * kFastThreads * 4 + (size_t)(q8 + kSlotPad) * kFastThreads + smem_pad;
  if (!fs.attr_set) {
    const int lim = 227 * 1024;
    cudaFuncSetAttribute(k_local_s16<START_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_POS>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_POS>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_NONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_POS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_NONE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_POS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_REV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_NONE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_POS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_NONE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_semiglobal_s16<START_POS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_global_s16, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_banded_tb_s16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_banded_tb_s16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_banded_tb_s16<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_banded_tb_s16<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_NONE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_POS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_NONE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_local_s16<START_POS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    fs.attr_set = true;
  }
  k_zero_u32<<<1, 1, 0, str>>>(fs.work_count);
  int launches = 2;
  if (passes) {
    // which pass scores which pair (one coalesced read of the packed sequences); SEMI_GLOBAL / GLOBAL cannot run empty sequences
    cudaMemsetAsync(fs.flags, 0, sizeof(uint32_t) * (size_t)a.n, str);
    const uint32_t chunks = std::max<uint32_t>(1u, (((max_q + 7u) >> 3) + ((max_t + 7u) >> 3) + 7u) / 8u);
    const uint64_t threads = (uint64_t)a.n * chunks;
    k_classify<<<(uint32_t)((threads + 255) / 256), 256, 0, str>>>(fa, fs.flags, chunks, nmask ? 1 : 0, algo != ALGO_LOCAL, algo == ALGO_GLOBAL);
    launches += 2;
  }
  const uint32_t blocks = ((uint32_t)a.n + kFastThreads - 1) / kFastThreads;
  const bool sp = (start == START_POS);
  const bool tbm = (start == START_TB);
  const bool tailq = (algo == ALGO_SEMI_GLOBAL) && (a.tail == SKIP_QUERY || a.tail == SKIP_BOTH);
  // grouped arg-max (keys score * 32 + ...) while every possible score stays below 1024, else the per-step keys (score * 8 + ...)
  const uint32_t t8 = (max_t + 7u) & ~7u;
  const bool g4 = GB200_LOCAL_G4 && (uint64_t)std::min(q8, t8) * (uint64_t)a.match <= 1000;
  if (nmask && passes) {
    // N-aware launch on the side stream, beside the main launch.  Its grid covers the worst case (every pair flagged); CTAs
    // beyond work_count[0] exit on their first instruction.
    cudaEventRecord(fs.ev_fork, str);
    cudaStreamWaitEvent(fs.side, fs.ev_fork, 0);
    if (banded_sc) k_banded_tb_s16<true, false><<<blocks, kFastThreads, smem, fs.side>>>(fa);
    else if (banded) k_banded_tb_s16<true><<<blocks, kFastThreads, smem, fs.side>>>(fa);
    else if (algo == ALGO_LOCAL) {
      if (tbm) k_local_s16<START_TB, true><<<blocks, kFastThreads, smem, fs.side>>>(fa);
      else if (g4) { if (sp) k_local_s16<START_POS, true, true><<<blocks, kFastThreads, smem, fs.side>>>(fa); else k_local_s16<START_NONE, true, true><<<blocks, kFastThreads, smem, fs.side>>>(fa); }
      else if (sp) k_local_s16<START_POS, true><<<blocks, kFastThreads, smem, fs.side>>>(fa); else k_local_s16<START_NONE, true><<<blocks, kFastThreads, smem, fs.side>>>(fa);
    } else if (tailq) { if (sp) k_semiglobal_s16<START_POS, true, true><<<blocks, kFastThreads, smem, fs.side>>>(fa); else k_semiglobal_s16<START_NONE, true, true><<<blocks, kFastThreads, smem, fs.side>>>(fa); }
    else if (sp && GB200_SEMI_SPLIT_START) {
      k_semiglobal_s16<START_NONE, true><<<blocks, kFastThreads, smem, fs.side>>>(fa);
      k_semiglobal_s16<START_REV, true><<<blocks, kFastThreads, smem, fs.side>>>(fa);
      launches++;
    }
    else { if (sp) k_semiglobal_s16<START_POS, true><<<blocks, kFastThreads, smem, fs.side>>>(fa); else k_semiglobal_s16<START_NONE, true><<<blocks, kFastThreads, smem, fs.side>>>(fa); }
    cudaEventRecord(fs.ev_join, fs.side);
    launches++;
  }
  if (banded_sc) k_banded_tb_s16<false, false><<<blocks, kFastThreads, smem, str>>>(fa);
  else if (banded) k_banded_tb_s16<false><<<blocks, kFastThreads, smem, str>>>(fa);
  else if (algo == ALGO_LOCAL) {
    if (tbm) k_local_s16<START_TB><<<blocks, kFastThreads, smem, str>>>(fa);
    else if (g4) { if (sp) k_local_s16<START_POS, false, true><<<blocks, kFastThreads, smem, str>>>(fa); else k_local_s16<START_NONE, false, true><<<blocks, kFastThreads, smem, str>>>(fa); }
    else if (sp) k_local_s16<START_POS><<<blocks, kFastThreads, smem, str>>>(fa); else k_local_s16<START_NONE><<<blocks, kFastThreads, smem, str>>>(fa);
  }
  else if (algo == ALGO_SEMI_GLOBAL && tailq) { if (sp) k_semiglobal_s16<START_POS, false, true><<<blocks, kFastThreads, smem, str>>>(fa); else k_semiglobal_s16<START_NONE, false, true><<<blocks, kFastThreads, smem, str>>>(fa); }
  else if (algo == ALGO_SEMI_GLOBAL && sp && GB200_SEMI_SPLIT_START) {
    k_semiglobal_s16<START_NONE><<<blocks, kFastThreads, smem, str>>>(fa);
    k_semiglobal_s16<START_REV><<<blocks, kFastThreads, smem, str>>>(fa);
    launches++;
  }
  else if (algo == ALGO_SEMI_GLOBAL) { if (sp) k_semiglobal_s16<START_POS><<<blocks, kFastThreads, smem, str>>>(fa); else k_semiglobal_s16<START_NONE><<<blocks, kFastThreads, smem, str>>>(fa); }
  else k_global_s16<<<blocks, kFastThreads, smem, str>>>(fa);
  if (nmask && passes) cudaStreamWaitEvent(str, fs.ev_join, 0);
  if (!passes) return launches;             // (the kernels' own safety-net lists stay empty: the packer's verdict is exact)
  // general kernel over what is left, bounded scratch
  const int which = 1;                      // k_classify files everything the s16x2 passes cannot score in worklist2
  fa.a.colstate = fs.colstate; fa.a.col_stride = slots; fa.a.col_rows = q8;
  const int wb = (int)(slots / 128);
  if (banded_sc) k_worklist<ALGO_BANDED, START_NONE><<<wb, 128, 0, str>>>(fa, which);
  else if (algo == ALGO_LOCAL) {
    if (tbm) k_worklist<ALGO_LOCAL, START_TB><<<wb, 128, 0, str>>>(fa, which);
    else if (sp) k_worklist<ALGO_LOCAL, START_POS><<<wb, 128, 0, str>>>(fa, which); else k_worklist<ALGO_LOCAL, START_NONE><<<wb, 128, 0, str>>>(fa, which);
  }
  else if (algo == ALGO_SEMI_GLOBAL) { if (sp) k_worklist<ALGO_SEMI_GLOBAL, START_POS><<<wb, 128, 0, str>>>(fa, which); else k_worklist<ALGO_SEMI_GLOBAL, START_NONE><<<wb, 128, 0, str>>>(fa, which); }
  else k_worklist<ALGO_GLOBAL, START_NONE><<<wb, 128, 0, str>>>(fa, which);
  return launches;
}

}  // namespace gb200
