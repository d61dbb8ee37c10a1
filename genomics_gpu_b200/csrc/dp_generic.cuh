// dp_generic.cuh -- general inter-task DP kernels: one thread per pair, int32 lanes on the
// Blackwell DPX instructions (__vimax3_s32[_relu] -> VIMNMX3[.RELU], __viaddmax_s32 -> VIADDMNMX).
//
// These kernels cover EVERY mode of the reference bit-exactly and with no length limit:
//   local / global / semi-global (16 head/tail variants) / banded, score-only, WITH_START
//   (reverse pass), second-best, and the 4-bit direction matrix for traceback.
// They are the exact fallback behind the s16x2 fast path (dp_s16.cuh), which handles the common
// short-read shapes.  What each mode computes, including the reference's quirks, is documented in
// SURVEY.md section 8 (a11-a15, Q1-Q14); citations below are relative to
// /root/reference/Non-CDP/GASAL2/src/kernels/.
//
// Differences in HOW (not what) vs the reference:
//   * the per-query-column (H,E) carry lives in a coalesced global scratch [column][pair] that is
//     sized at run time -- no MAX_QUERY_LEN, no per-thread local-memory array, no O(MAX_QUERY_LEN)
//     init loop per thread (only the pair's own columns are initialised);
//   * the running arg-max is tracked per COLUMN (one max per cell, position resolved only when a
//     column really improves the best), which yields the same (strip, column, row) winner;
//   * scoring constants are kernel arguments; N_CODE / N_PENALTY are run-time;
//   * the reversed sequences of the semi-global start pass are fetched on the fly.
#pragma once
#include "dp_common.cuh"

namespace gb200 {

struct PairCtx {
  const uint32_t* qw; const uint32_t* tw;
  uint32_t qlen, tlen, QR, TR;
  uint32_t* cs;       // this pair's column-state base (index with c * stride)
  uint32_t stride;
};

struct Sc { int match, nmis, oe, ext, o; uint32_t nval; int nscore; };

__device__ __forceinline__ Sc make_sc(const DpArgs& a, bool global_rule) {
  Sc s; s.match = a.match; s.nmis = -a.mismatch; s.o = a.gap_open; s.ext = a.gap_extend; s.oe = a.gap_open + a.gap_extend;
  // GLOBAL rule: N is an ordinary symbol unless N_PENALTY is defined (gasal_kernels.h:39-56)
  s.nval = (global_rule && !a.n_is_wild_global) ? 0xFFFFFFFFu : a.n_value;
  s.nscore = a.n_score;
  return s;
}

// Per-strip view of the 8 target bases: tg[r] = code (or an impossible code if it is N),
// tmm[r] = score of a non-match in that row (mismatch penalty, or the N score).
struct StripRows { uint32_t g[8]; int mm[8]; };

__device__ __forceinline__ void load_rows(const Sc& sc, uint32_t gpac, StripRows& R) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const uint32_t g = nib(gpac, r);
    const bool isn = (g == sc.nval);
    R.g[r] = isn ? 0x100u : g;
    R.mm[r] = isn ? sc.nscore : sc.nmis;
  }
}

// ------------------------------------------------------------------------------------------------
// LOCAL forward pass (local_kernel_template.h:71-438).  TB: also emit direction nibbles.
// SECOND: literal per-cell second-best bookkeeping (:147-163).
template <bool TB, bool SECOND>
__device__ __forceinline__ void local_forward(const DpArgs& a, const Sc& sc, const PairCtx& p, uint32_t tid,
                                              int& best, int& best_q, int& best_t) {
  best = 0; best_q = 0; best_t = 0;
  int prev_best = 0;
  int best2 = 0, prev_best2 = 0, best2_q = 0, best2_t = 0;
  const uint32_t ncol = p.QR * 8;
  for (uint32_t c = 0; c < ncol; c++) p.cs[(size_t)c * p.stride] = 0u;               // :118-120
  uint32_t tile_no = 0;
  for (uint32_t i = 0; i < p.TR; i++) {
    int diag[8], F[8];
#pragma unroll
    for (int r = 0; r < 8; r++) { diag[r] = 0; F[r] = 0; }                              // :123-127
    StripRows R; load_rows(sc, p.tw[i], R);
    uint32_t c = 0;
    for (uint32_t j = 0; j < p.QR; j++) {
      const uint32_t rpac = p.qw[j];
      uint4 dir = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; k++, c++) {
        const uint32_t rb = nib(rpac, k);
        const bool rn = (rb == sc.nval);
        const short2 hd = u32_as_short2(p.cs[(size_t)c * p.stride]);                    // :390-392
        int up = hd.x, e = hd.y;
        int colmax = 0;
        int hcol[8];
        uint32_t dw = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {                                                   // CORE_LOCAL_COMPUTE :19-30
          int s = (rb == R.g[r]) ? sc.match : R.mm[r];
          s = rn ? sc.nscore : s;
          const int d = diag[r] + s;
          const int h = __vimax3_s32_relu(d, F[r], e);
          const int dm = d - sc.oe;
          if (TB) {                                                                     // :45-60
            const uint32_t mx = (s >= 0) ? 0u : 1u;
            const uint32_t code = (h == d) ? mx : ((h == F[r]) ? 3u : 2u);
            const uint32_t fb = (dm > F[r] - sc.ext) ? 0u : 8u;
            const uint32_t eb = (dm > e - sc.ext) ? 0u : 4u;
            dw |= (code | fb | eb) << (28 - 4 * r);
          }
          F[r] = __viaddmax_s32(F[r], -sc.ext, dm);
          e = __viaddmax_s32(e, -sc.ext, dm);
          if (SECOND) {
            if (best < h) { best_t = (int)(8 * i + r); best = h; }
            if (best2 < h && best > h) { best2_t = (int)(8 * i + r); best2 = h; }
          } else {
            hcol[r] = h;
            colmax = max(colmax, h);
          }
          diag[r] = up; up = h;
        }
        p.cs[(size_t)c * p.stride] = pack_short2(up, e);                                // :406-408
        if (TB) {
          if ((k & 3) == 0) dir.x = dw; else if ((k & 3) == 1) dir.y = dw; else if ((k & 3) == 2) dir.z = dw; else dir.w = dw;
          if ((k & 3) == 3) { a.tb[(size_t)tile_no * a.tb_stride + tid] = dir; tile_no++; dir = make_uint4(0, 0, 0, 0); }   // :257,380
        }
        if (SECOND) {
          if (prev_best < best) best_q = (int)c;
          if (prev_best2 < best) best2_q = (int)c;                                      // (sic) :162
          prev_best2 = max(best2, prev_best2);
          prev_best = max(best, prev_best);
        } else if (colmax > best) {
          // first row (in row order) that reaches the column maximum == the reference's last update
          int rr = 7;
#pragma unroll
          for (int r = 6; r >= 0; r--) rr = (hcol[r] == colmax) ? r : rr;
          best = colmax; best_q = (int)c; best_t = (int)(8 * i) + rr;
        }
      }
    }
  }
  if (SECOND) {
    if (a.score2) a.score2[tid] = best2;
    if (a.q_end2) a.q_end2[tid] = best2_q;
    if (a.t_end2) a.t_end2[tid] = best2_t;
  }
}

// LOCAL reverse pass for the start position (:441-511).  Walks the words that hold the end cell
// backwards; stops as soon as the running max reaches the forward score (checked between columns);
// target_batch_start keeps the reference's mirrored-in-word value.
__device__ __forceinline__ void local_reverse(const Sc& sc, const PairCtx& p, int fwd, int end_q, int end_t, int& sx, int& sy) {
  const int rend_reg = min((end_q >> 3) + 1, (int)p.QR);
  const int gend_reg = min((end_t >> 3) + 1, (int)p.TR);
  int m = 0; sx = 0; sy = 0;
  const int ncol = rend_reg * 8;
  for (int c = 0; c < ncol; c++) p.cs[(size_t)c * p.stride] = 0u;
  int gidx = ((gend_reg << 3) + 8) - 1;
  for (int i = 0; i < gend_reg && m < fwd; i++) {
    int diag[8], F[8];
#pragma unroll
    for (int r = 0; r < 8; r++) { diag[r] = 0; F[r] = 0; }
    StripRows R; load_rows(sc, nibrev(p.tw[gend_reg - 1 - i]), R);
    gidx -= 8;
    int ridx = (rend_reg << 3) - 1;
    int c = 0;
    for (int j = 0; j < rend_reg && m < fwd; j++) {
      const uint32_t rpac = nibrev(p.qw[rend_reg - 1 - j]);
      for (int k = 0; k < 8 && m < fwd; k++, c++, ridx--) {
        const uint32_t rb = nib(rpac, k);
        const bool rn = (rb == sc.nval);
        const short2 hd = u32_as_short2(p.cs[(size_t)c * p.stride]);
        int up = hd.x, e = hd.y;
        int colmax = 0; int hcol[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
          int s = (rb == R.g[r]) ? sc.match : R.mm[r];
          s = rn ? sc.nscore : s;
          const int d = diag[r] + s;
          const int h = __vimax3_s32_relu(d, F[r], e);
          const int dm = d - sc.oe;
          F[r] = __viaddmax_s32(F[r], -sc.ext, dm);
          e = __viaddmax_s32(e, -sc.ext, dm);
          hcol[r] = h; colmax = max(colmax, h);
          diag[r] = up; up = h;
        }
        p.cs[(size_t)c * p.stride] = pack_short2(up, e);
        if (colmax > m) {
          int rr = 7;
#pragma unroll
          for (int r = 6; r >= 0; r--) rr = (hcol[r] == colmax) ? r : rr;
          m = colmax; sx = ridx; sy = gidx + rr;                                        // Q3: gidx + (m-1)
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SEMI-GLOBAL (semiglobal_kernel_template.h:39-388).  One pass; REV selects the reversed sequences
// of the WITH_START pass.  Tracks (best, x, y) exactly as the reference's (maxHH, maxXY_x, maxXY_y).
struct SgTrack { int best, x, y, best2, x2, y2; };

template <bool REV, bool SECOND>
__device__ __forceinline__ void semiglobal_pass(const DpArgs& a, const Sc& sc, const PairCtx& p, uint32_t first_strip,
                                                int stop_score, SgTrack& tr) {
  const bool head_q = (a.head == SKIP_QUERY || a.head == SKIP_BOTH);                    // top row free   :87
  const bool head_t = (a.head == SKIP_TARGET || a.head == SKIP_BOTH);                   // left col free  :114
  const bool tail_t = (a.tail == SKIP_TARGET || a.tail == SKIP_BOTH);
  const bool tail_q = (a.tail == SKIP_QUERY || a.tail == SKIP_BOTH);
  const uint32_t ncol = p.QR * 8;
  if (head_q) { for (uint32_t c = 0; c < ncol; c++) p.cs[(size_t)c * p.stride] = 0u; }
  else {
    p.cs[0] = pack_short2(0, kMinusInf);
    for (uint32_t c = 1; c < ncol; c++) p.cs[(size_t)c * p.stride] = pack_short2(-(sc.o + sc.ext * (int)c), kMinusInf);   // :94-98
  }
  int row_ctr = 0;
  for (uint32_t i = first_strip; i < p.TR && (!REV || tr.best < stop_score); i++) {
    int Hl[8], F[8], diag[8];
    if (head_t) {
#pragma unroll
      for (int r = 0; r < 8; r++) { Hl[r] = 0; F[r] = kMinusInf; diag[r] = 0; }
    } else {
#pragma unroll
      for (int r = 0; r < 8; r++, row_ctr++) {
        Hl[r] = -(sc.o + sc.ext * row_ctr); F[r] = kMinusInf;                           // :123-128
        diag[r] = (row_ctr == 0) ? 0 : -(sc.o + sc.ext * row_ctr);
      }
    }
    StripRows R; load_rows(sc, REV ? rev_true_word(p.tw, (int)p.tlen, (int)i) : p.tw[i], R);
    if (REV) {
      // zero pad nibbles of the reversed copies are ordinary symbols: 0 == 0 matches
    }
    uint32_t c = 0;
    for (uint32_t j = 0; j < p.QR && (!REV || tr.best < stop_score); j++) {
      const uint32_t rpac = REV ? rev_true_word(p.qw, (int)p.qlen, (int)j) : p.qw[j];
#pragma unroll
      for (int k = 0; k < 8; k++, c++) {
        const uint32_t rb = nib(rpac, k);
        const bool rn = (rb == sc.nval);
        const short2 hd = u32_as_short2(p.cs[(size_t)c * p.stride]);
        int up = hd.x, e = hd.y;
        int hcol[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {                                                   // CORE_COMPUTE_SEMIGLOBAL :17-28
          int s = (rb == R.g[r]) ? sc.match : R.mm[r];
          s = rn ? sc.nscore : s;
          F[r] = __viaddmax_s32(F[r], -sc.ext, Hl[r] - sc.oe);
          e = __viaddmax_s32(e, -sc.ext, up - sc.oe);
          const int h = __vimax3_s32(diag[r] + s, F[r], e);
          Hl[r] = h; hcol[r] = h; diag[r] = up; up = h;
        }
        p.cs[(size_t)c * p.stride] = pack_short2(up, e);
        if (tail_t && c + 1 == p.qlen) {                                                // :160-178
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int t = (int)(8 * i + r);
            const bool in = (uint32_t)t < p.tlen;
            if (hcol[r] > tr.best && in) { tr.y = t; tr.best = hcol[r]; }
            if (SECOND && hcol[r] > tr.best2 && hcol[r] < tr.best && in) { tr.y2 = t; tr.best2 = hcol[r]; }
          }
        }
      }
    }
  }
  if (tail_q) {                                                                         // :185-216
    // The reference scans global[0 .. MAX_QUERY_LEN): beyond this pair's columns the array still
    // holds its initial values, which only the (buggy) second-best test `m < ref_len` can reach.
    const uint32_t lim = SECOND ? max(p.qlen, p.tlen) : p.qlen;
    for (uint32_t m = 0; m < lim; m++) {
      int v;
      if (m < ncol) v = u32_as_short2(p.cs[(size_t)m * p.stride]).x;
      else v = head_q ? 0 : (int)(short)(-(sc.o + sc.ext * (int)m));
      if (v > tr.best && m < p.qlen) { tr.x = (int)m; tr.best = v; }
      if (SECOND && v > tr.best2 && v < tr.best && m < p.tlen) { tr.x2 = (int)m; tr.best2 = v; }
    }
    if (tr.x != (int)p.tlen) tr.y = (int)p.qlen;
    if (SECOND && tr.x2 != (int)p.tlen) tr.y2 = (int)p.qlen;
  }
}

// ------------------------------------------------------------------------------------------------
// GLOBAL (global.h:30-303): D-based opening, no clamp, shifted boundaries (Q4); score only, or TB.
template <bool TB>
__device__ __forceinline__ int global_forward(const DpArgs& a, const Sc& sc, const PairCtx& p, uint32_t tid) {
  const uint32_t ncol = p.QR * 8;
  p.cs[0] = pack_short2(0, kMinusInf);                                                  // :57-60
  for (uint32_t c = 1; c < ncol; c++) p.cs[(size_t)c * p.stride] = pack_short2(-(sc.o + sc.ext * (int)c), kMinusInf);
  int last[8];
#pragma unroll
  for (int r = 0; r < 8; r++) last[r] = 0;
  int row = 0;
  uint32_t tile_no = 0;
  for (uint32_t i = 0; i < p.TR; i++) {
    int diag[8], F[8];
#pragma unroll
    for (int r = 0; r < 8; r++, row++) { F[r] = kMinusInf; diag[r] = (row == 0) ? 0 : -(sc.o + sc.ext * row); }   // :66-70
    StripRows R; load_rows(sc, p.tw[i], R);
    uint32_t c = 0;
    for (uint32_t j = 0; j < p.QR; j++) {
      const uint32_t rpac = p.qw[j];
      uint4 dir = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; k++, c++) {
        const uint32_t rb = nib(rpac, k);
        const bool rn = (rb == sc.nval);
        const short2 hd = u32_as_short2(p.cs[(size_t)c * p.stride]);
        int up = hd.x, e = hd.y;
        int hcol[8];
        uint32_t dw = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {                                                   // CORE_GLOBAL_COMPUTE :4-12
          int s = (rb == R.g[r]) ? sc.match : R.mm[r];
          s = rn ? sc.nscore : s;
          const int d = diag[r] + s;
          const int h = __vimax3_s32(d, F[r], e);
          const int dm = d - sc.oe;
          if (TB) {                                                                     // :14-26
            const uint32_t mx = (s >= 0) ? 0u : 1u;
            const uint32_t code = (h == d) ? mx : ((h == F[r]) ? 3u : 2u);
            const uint32_t fb = (dm > F[r] - sc.ext) ? 0u : 8u;
            const uint32_t eb = (dm > e - sc.ext) ? 0u : 4u;
            dw |= (code | fb | eb) << (28 - 4 * r);
          }
          F[r] = __viaddmax_s32(F[r], -sc.ext, dm);
          e = __viaddmax_s32(e, -sc.ext, dm);
          hcol[r] = h; diag[r] = up; up = h;
        }
        p.cs[(size_t)c * p.stride] = pack_short2(up, e);
        if (TB) {
          if ((k & 3) == 0) dir.x = dw; else if ((k & 3) == 1) dir.y = dw; else if ((k & 3) == 2) dir.z = dw; else dir.w = dw;
          if ((k & 3) == 3) { a.tb[(size_t)tile_no * a.tb_stride + tid] = dir; tile_no++; dir = make_uint4(0, 0, 0, 0); }
        }
        if (c + 1 == p.qlen) {                                                          // :285-290
#pragma unroll
          for (int r = 0; r < 8; r++) last[r] = hcol[r];
        }
      }
    }
  }
  const int idx = 7 - (int)(p.TR * 8 - p.tlen);                                         // :299
  int v = last[0];
#pragma unroll
  for (int r = 1; r < 8; r++) v = (idx == r) ? last[r] : v;
  return v;
}

// ------------------------------------------------------------------------------------------------
// BANDED (banded.h:10-139): local-type, H-based opening, band in whole 8x8 tiles.
__device__ __forceinline__ void banded_forward(const DpArgs& a, const Sc& sc, const PairCtx& p, int& best, int& best_q, int& best_t) {
  const int QR = (int)p.QR, TR = (int)p.TR;
  const int kw = a.k_band >> 3;                                                         // gasal_align.h:80
  const int k_other = TR - (QR - kw);                                                   // :35
  best = 0; best_q = 0; best_t = 0;
  for (int c = 0; c < QR * 8; c++) p.cs[(size_t)c * p.stride] = 0u;
  for (int i = 0; i < TR; i++) {
    int Hl[8], F[8], diag[8];
#pragma unroll
    for (int r = 0; r < 8; r++) { Hl[r] = 0; F[r] = 0; diag[r] = 0; }
    StripRows R; load_rows(sc, p.tw[i], R);
    int c = max(0, i - k_other + 1) << 3;                                               // :83
    const int last_tile = min(kw + i, QR);                                              // :84
    for (int j = c >> 3; j < last_tile; j++) {
      const uint32_t rpac = p.qw[j];
#pragma unroll
      for (int k = 0; k < 8; k++, c++) {
        const uint32_t rb = nib(rpac, k);
        const bool rn = (rb == sc.nval);
        const short2 hd = u32_as_short2(p.cs[(size_t)c * p.stride]);
        int up = hd.x, e = hd.y;
        int colmax = 0; int hcol[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {                                                   // :99-114
          int s = (rb == R.g[r]) ? sc.match : R.mm[r];
          s = rn ? sc.nscore : s;
          F[r] = __viaddmax_s32(F[r], -sc.ext, Hl[r] - sc.oe);
          e = __viaddmax_s32(e, -sc.ext, up - sc.oe);
          int h = __vimax3_s32_relu(diag[r] + s, F[r], e);
          Hl[r] = h; hcol[r] = h; colmax = max(colmax, h);
          diag[r] = up; up = h;
        }
        p.cs[(size_t)c * p.stride] = pack_short2(up, e);
        if (colmax > best) {
          int rr = 7;
#pragma unroll
          for (int r = 6; r >= 0; r--) rr = (hcol[r] == colmax) ? r : rr;
          best = colmax; best_q = c; best_t = 8 * i + rr;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BANDED + WITH_TB (extension; the reference has no banded traceback -- SURVEY 8c, config 4): the LOCAL recurrence and
// direction nibbles (local_kernel_template.h:19-60) over the tile band of banded.h:35,83-85, cells outside the band read
// as 0 exactly as in the banded kernel; only the band's tiles are written (0.5 B per in-band cell).
__device__ __forceinline__ void banded_tb_forward(const DpArgs& a, const Sc& sc, const PairCtx& p, uint32_t tid, int& best, int& best_q, int& best_t) {
  const int QR = (int)p.QR, TR = (int)p.TR;
  const BandGeom bg = band_geom(QR, TR, a.k_band);
  best = 0; best_q = 0; best_t = 0;
  for (int c = 0; c < QR * 8; c++) p.cs[(size_t)c * p.stride] = 0u;
  for (int i = 0; i < TR; i++) {
    int F[8], diag[8];
#pragma unroll
    for (int r = 0; r < 8; r++) { F[r] = 0; diag[r] = 0; }
    StripRows R; load_rows(sc, p.tw[i], R);
    const int lo = band_lo(bg, i), hi = band_hi(bg, i, QR);
    int c = lo << 3;
    for (int j = lo; j < hi; j++) {
      const uint32_t rpac = p.qw[j];
      uint4 dir = make_uint4(0, 0, 0, 0);
      const size_t tile0 = ((size_t)i * bg.W + (j - lo)) * 2;
#pragma unroll
      for (int k = 0; k < 8; k++, c++) {
        const uint32_t rb = nib(rpac, k);
        const bool rn = (rb == sc.nval);
        const short2 hd = u32_as_short2(p.cs[(size_t)c * p.stride]);
        int up = hd.x, e = hd.y;
        int colmax = 0; int hcol[8];
        uint32_t dw = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          int s = (rb == R.g[r]) ? sc.match : R.mm[r];
          s = rn ? sc.nscore : s;
          const int d = diag[r] + s;
          const int h = __vimax3_s32_relu(d, F[r], e);
          const int dm = d - sc.oe;
          const uint32_t mx = (s >= 0) ? 0u : 1u;
          const uint32_t code = (h == d) ? mx : ((h == F[r]) ? 3u : 2u);
          const uint32_t fb = (dm > F[r] - sc.ext) ? 0u : 8u;
          const uint32_t eb = (dm > e - sc.ext) ? 0u : 4u;
          dw |= (code | fb | eb) << (28 - 4 * r);
          F[r] = __viaddmax_s32(F[r], -sc.ext, dm);
          e = __viaddmax_s32(e, -sc.ext, dm);
          hcol[r] = h; colmax = max(colmax, h);
          diag[r] = up; up = h;
        }
        p.cs[(size_t)c * p.stride] = pack_short2(up, e);
        if ((k & 3) == 0) dir.x = dw; else if ((k & 3) == 1) dir.y = dw; else if ((k & 3) == 2) dir.z = dw; else dir.w = dw;
        if ((k & 3) == 3) { a.tb[(tile0 + (k >> 2)) * a.tb_stride + tid] = dir; dir = make_uint4(0, 0, 0, 0); }
        if (colmax > best) {
          int rr = 7;
#pragma unroll
          for (int r = 6; r >= 0; r--) rr = (hcol[r] == colmax) ? r : rr;
          best = colmax; best_q = c; best_t = 8 * i + rr;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// KSW seed extension (ksw_kernel_template.h:46-199 = BWA's ksw_extend with z-drop off): row-major sweep with the adaptive
// [beg, end) column window, initial score h0 from seed_scores, the 5-point clipping rule at the end.  One thread per
// pair; the (H, E) row lives in the coalesced scratch as int32 pairs (the reference keeps eh_t eh[MAX_QUERY_LEN] per
// thread in local memory and zeroes all of it).  Quirk kept: the query word count is (qlen >> 3) + 1, so when qlen is a
// multiple of 8 the loop variable ends on j == qlen in EVERY row and the to-end score is always considered (:57,142-145).
// Where the (H, E) row of a pair lives.  KswRowGlobal: int32 pairs in the coalesced scratch ([column][pair]; any query length, any
// score).  KswRowShared: u16 pairs in shared memory ([column][thread of the CTA]) -- every stored value is >= 0 (H and E are clamped
// at 0, ksw_kernel_template.h:126-137), so 16 unsigned bits hold it while h0 + qlen * match <= 65535 (host check): the row never
// leaves the SM, where the global row costs 16 B of L2 / HBM traffic per cell (126 GB for 262144 extensions of 150 x 200).
struct KswRowGlobal {
  uint2* eh; size_t st;
  __device__ __forceinline__ void get(int j, int& h, int& e) const { const uint2 v = eh[(size_t)j * st]; h = (int)v.x; e = (int)v.y; }
  __device__ __forceinline__ void set(int j, int h, int e) const { eh[(size_t)j * st] = make_uint2((uint32_t)h, (uint32_t)e); }
  __device__ __forceinline__ bool nonzero(int j) const { const uint2 v = eh[(size_t)j * st]; return v.x != 0u || v.y != 0u; }
};
template <int THREADS>
struct KswRowShared {
  uint32_t* eh;                                               // this thread's column 0
  __device__ __forceinline__ void get(int j, int& h, int& e) const { const uint32_t v = eh[j * THREADS]; h = (int)(v & 0xFFFFu); e = (int)(v >> 16); }
  __device__ __forceinline__ void set(int j, int h, int e) const { eh[j * THREADS] = (uint32_t)h | ((uint32_t)e << 16); }
  __device__ __forceinline__ bool nonzero(int j) const { return eh[j * THREADS] != 0u; }
};

template <typename Row>
__device__ __forceinline__ void ksw_extend(const DpArgs& a, const Sc& sc, const PairCtx& p, uint32_t tid, const Row eh, int& o_score, int& o_qe, int& o_te) {
  const int qlen = (int)p.qlen, tlen = (int)p.tlen, h0 = (int)a.seed_scores[tid];
  const int e_del = sc.ext, e_ins = sc.ext, oe_del = sc.oe, oe_ins = sc.oe;
  for (int j = 0; j <= qlen + 1; j++) eh.set(j, 0, 0);
  eh.set(0, h0, 0);
  int prev = h0 > oe_ins ? h0 - oe_ins : 0;
  eh.set(1, prev, 0);
  for (int j = 2; j <= qlen && prev > e_ins; ++j) { prev -= e_ins; eh.set(j, prev, 0); }
  int mx = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, beg = 0, end = qlen;
  const int qt_last = qlen >> 3;                              // last query word the reference visits
  bool stop = false;
  for (int i = 0; i < tlen && !stop; i++) {
    const uint32_t gbase = nib(p.tw[i >> 3], i & 7);
    const bool gn = (gbase == sc.nval);
    int f = 0, h1, m = 0, mj = -1;
    if (beg == 0) { h1 = h0 - (sc.o + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
    uint32_t rpac = 0;
    for (int j = beg; j < end; j++) {
      if ((j & 7) == 0 || j == beg) rpac = p.qw[j >> 3];
      const uint32_t rbase = nib(rpac, j & 7);
      int M, e; eh.get(j, M, e);
      int s = (rbase == gbase) ? sc.match : sc.nmis;
      s = (gn || rbase == sc.nval) ? sc.nscore : s;
      M = M ? M + s : 0;                                      // separating H and M to disallow a cigar like "100M3I3D20M" (:126)
      int h = M > e ? M : e; h = h > f ? h : f;
      mj = m > h ? mj : j;
      m = m > h ? m : h;
      int t = M - oe_del; t = t > 0 ? t : 0; e -= e_del; e = e > t ? e : t;
      eh.set(j, h1, e);                                       // H(i,j-1) for the next row, E(i+1,j)
      h1 = h;
      t = M - oe_ins; t = t > 0 ? t : 0; f -= e_ins; f = f > t ? f : t;
    }
    eh.set(end, h1, 0);
    const int j_final = max(end, qt_last * 8);                // value of the reference's loop variable after the sweep
    if (j_final == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
    if (m == 0) {
      // `break` only leaves the reference's 8-row loop (:150): the sweep resumes with the next target word
      i |= 7;
      continue;
    }
    if (m > mx) { mx = m; max_i = i; max_j = mj; }
    int j = beg;
    for (; j < end; ++j) if (eh.nonzero(j)) break;
    beg = j;
    for (j = end; j >= beg; --j) if (eh.nonzero(j)) break;
    end = j + 2 < qlen ? j + 2 : qlen;
  }
  if (gscore <= 0 || gscore <= mx - 5) { o_score = mx; o_qe = max_j + 1; o_te = max_i + 1; }    // PEN_CLIP5 (:22,189-198)
  else { o_score = gscore; o_qe = qlen; o_te = max_ie + 1; }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool make_ctx(const DpArgs& a, uint32_t tid, PairCtx& p) {
  if (tid >= (uint32_t)a.n) return false;
  p.qw = a.pq + (a.q_offs[tid] >> 3);
  p.tw = a.pt + (a.t_offs[tid] >> 3);
  p.qlen = a.q_lens[tid]; p.tlen = a.t_lens[tid];
  p.QR = n_words(p.qlen); p.TR = n_words(p.tlen);
  p.cs = a.colstate + tid; p.stride = a.col_stride;
  return true;
}

template <int START, bool SECOND>
__global__ void __launch_bounds__(128) k_local_generic(const DpArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  PairCtx p; if (!make_ctx(a, tid, p)) return;
  const Sc sc = make_sc(a, false);
  int best, bq, bt;
  local_forward<START == START_TB, SECOND>(a, sc, p, tid, best, bq, bt);
  a.score[tid] = best; a.q_end[tid] = bq; a.t_end[tid] = bt;
  if (START == START_POS) {
    int sx, sy; local_reverse(sc, p, best, bq, bt, sx, sy);
    a.q_start[tid] = sx; a.t_start[tid] = sy;
  }
}

// one pair, forward (+ start pass); `pi` indexes the result arrays
template <int START, bool SECOND>
__device__ __forceinline__ void semiglobal_pair(const DpArgs& a, const Sc& sc, const PairCtx& p, uint32_t pi) {
  SgTrack tr; tr.best = kMinusInf; tr.x = (int)p.tlen; tr.y = (int)p.qlen;              // :63-64
  tr.best2 = kMinusInf; tr.x2 = (int)p.tlen; tr.y2 = (int)p.qlen;
  semiglobal_pass<false, SECOND>(a, sc, p, 0, 0, tr);
  a.score[pi] = tr.best; a.t_end[pi] = tr.y; a.q_end[pi] = tr.x;
  if (SECOND) { a.score2[pi] = tr.best2; a.t_end2[pi] = tr.y2; a.q_end2[pi] = tr.x2; }
  if (START == START_POS) {
    const int fwd = tr.best;
    // :273 -- unsigned arithmetic in the reference: a wrapped value makes the strip loop a no-op
    const uint32_t ua = p.TR - (uint32_t)((tr.y >> 3) + 1);
    const uint32_t gend_reg = ua > 0 ? ua - 1 : ua;
    tr.best = kMinusInf; tr.y = 0;                                                      // :275-276
    semiglobal_pass<true, false>(a, sc, p, gend_reg, fwd, tr);
    a.t_start[pi] = ((int)p.tlen - 1) - tr.y;                                           // :380-381
    a.q_start[pi] = ((int)p.qlen - 1) - tr.x;
  }
}

template <int START, bool SECOND>
__global__ void __launch_bounds__(128) k_semiglobal_generic(const DpArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  PairCtx p; if (!make_ctx(a, tid, p)) return;
  const Sc sc = make_sc(a, false);
  semiglobal_pair<START, SECOND>(a, sc, p, tid);
}

template <int START>
__global__ void __launch_bounds__(128) k_global_generic(const DpArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  PairCtx p; if (!make_ctx(a, tid, p)) return;
  const Sc sc = make_sc(a, true);
  a.score[tid] = global_forward<START == START_TB>(a, sc, p, tid);
}

static __global__ void __launch_bounds__(128) k_ksw_generic(const DpArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  PairCtx p; if (!make_ctx(a, tid, p)) return;
  p.cs = a.colstate + 2 * (size_t)tid;                        // int32 pairs: two words per column entry
  const Sc sc = make_sc(a, false);
  int score, qe, te;
  KswRowGlobal row; row.eh = (uint2*)p.cs; row.st = p.stride;
  ksw_extend(a, sc, p, tid, row, score, qe, te);
  a.score[tid] = score; a.q_end[tid] = qe; a.t_end[tid] = te;
}

// the same extension with the (H, E) row in shared memory (KswRowShared): dynamic shared memory = (max query length + 2) * 128 * 4 B
constexpr int kKswThreads = 128;
static __global__ void __launch_bounds__(kKswThreads) k_ksw_smem(const DpArgs a) {
  extern __shared__ uint32_t ksw_rows[];
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  PairCtx p; if (!make_ctx(a, tid, p)) return;
  const Sc sc = make_sc(a, false);
  int score, qe, te;
  KswRowShared<kKswThreads> row; row.eh = ksw_rows + threadIdx.x;
  ksw_extend(a, sc, p, tid, row, score, qe, te);
  a.score[tid] = score; a.q_end[tid] = qe; a.t_end[tid] = te;
}

template <bool TB>
__global__ void __launch_bounds__(128) k_banded_generic(const DpArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  PairCtx p; if (!make_ctx(a, tid, p)) return;
  const Sc sc = make_sc(a, false);
  int best, bq, bt;
  if (TB) banded_tb_forward(a, sc, p, tid, best, bq, bt); else banded_forward(a, sc, p, best, bq, bt);
  a.score[tid] = best; a.q_end[tid] = bq; a.t_end[tid] = bt;
}

}  // namespace gb200
