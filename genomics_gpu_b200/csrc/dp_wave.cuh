// dp_wave.cuh -- K3: the intra-task (warp-cooperative) anti-diagonal wavefront kernel for LONG pairs.
//
// One pair is cut into ROW-BLOCKS of 32 lanes x 8*SR target rows.  A row-block is one warp-task: lane L owns SR
// consecutive 8-row strips and walks the query left to right, one column per step, skewed by one column per lane
// (lane L is on column k - L at step k), so lane L's bottom row (H, E) reaches lane L+1 with __shfl_up_sync one step
// later -- no shared-memory column array at all inside a block.  Lane 31's bottom row is the top boundary of the next
// row-block of the same pair; it goes through a small per-warp shared-memory ring and is flushed 32 columns at a time
// (coalesced) to a per-pair double buffer in global memory (L2-resident).  Each 16-byte column entry is two
// self-validating 8-byte words (H, tag) (E, tag) -- the tag names the launch and the producing block -- so the consumer
// simply re-reads a chunk until every tag matches: no flags, no fences (8-byte stores are single-copy atomic; the same
// idea as NCCL's LL protocol), and the next chunk is prefetched a whole chunk ahead.  Tasks are handed out by a global
// ticket counter in BLOCK-MAJOR order (ticket = block * n_pairs + pair), so a task only ever waits on a LOWER ticket that
// some resident warp already holds: deadlock-free for any grid size, no cooperative launch.  With many pairs every worker
// starts on a different pair and a pair's next row-block is picked up long after its predecessor finished (no waiting at
// all); with few pairs the row-blocks of one pair run concurrently as a pipeline across the whole GPU.
//
// Same recurrences, boundaries and tie-breaking as the inter-task kernels (dp_generic.cuh), int32 lanes on the DPX
// instructions (VIMNMX3 / VIADDMNMX); the substitution score is ONE PRMT per cell from a per-row 4-byte table.
// Modes: LOCAL (score, end; start by a second launch over the reversed rectangle), GLOBAL (score), SEMI_GLOBAL with
// TAIL = TARGET and HEAD in {TARGET, BOTH} (score, end; start by a second launch over the true-length reversed
// sequences).  Everything else stays on the general inter-task kernels.
//
// Reference semantics: Non-CDP/GASAL2/src/kernels/local_kernel_template.h:71-519, global.h:30-303,
// semiglobal_kernel_template.h:39-388.  The reference has no intra-task kernel: one
// thread walks all 2.7e8 cells of a 16k x 16k pair with a 64 KB local-memory array (a13, "where the time goes").
#pragma once
#include "dp_common.cuh"
#include "dp_generic.cuh"
#include "dp_launch.h"

namespace gb200 {

constexpr int kWaveWarps = 4;              // independent warp-workers per CTA
constexpr uint32_t kWaveExotic = 1u << 16; // selector flag: query code outside the PRMT table -> compare path

struct WaveArgs {
  DpArgs a;
  int algo;                  // ALGO_LOCAL / ALGO_GLOBAL / ALGO_SEMI_GLOBAL
  int pass;                  // 0 = forward, 1 = start pass (reversed sequences; geometry from the forward results)
  uint32_t maxblk;           // tickets reserved per pair
  const uint64_t* qoff;      // per pair: first column slot in bnd (exclusive prefix sum of padded query lengths)
  const uint64_t* soff;      // per pair: first strip slot in cand (exclusive prefix sum of target words)
  uint4* bnd;                // two column buffers: bnd[buf * bnd_stride + qoff[p] + c] = (H, tag, E, tag) below the block's last row
  uint64_t bnd_stride;
  uint32_t tag_base;         // tag of block b = tag_base + b (unique per launch and block, never 0)
  int2* cand;                // per strip: (score, position) candidate of the strip
  uint32_t* ticket;          // global ticket counter
  uint32_t* done;            // per pair: finished row-blocks
  int kshift, floor16;       // k_wave16_global: constant subtracted from every biased value; finite -infinity (both fit 16 bits)
  uint32_t mm, sixteen;      // k_wave16_local_t<true>: 0x00204081 / 16 at run time (the direction-nibble gather stays on IMAD, see dp_s16.cuh)
};

struct WaveGeom {
  int ncol;        // query columns swept
  int nstrips;     // target strips swept (pass-local numbering)
  int strip0;      // first strip in the sequence numbering (semi-global start pass begins at gend_reg)
  int qwords, twords;   // local start pass: rend_reg, gend_reg
  int fwd;         // forward score (start passes)
};

// boundary entries travel through L2 only (never L1): relaxed gpu-scope vector accesses, each 8-byte half atomic
__device__ __forceinline__ uint4 ld_bnd(const uint4* p) {
  uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_bnd(uint4* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ WaveGeom wave_geometry(const WaveArgs& w, uint32_t pair) {
  const DpArgs& a = w.a;
  const int qlen = (int)a.q_lens[pair], tlen = (int)a.t_lens[pair];
  const int QR = (qlen + 7) >> 3, TR = (tlen + 7) >> 3;
  WaveGeom g; g.strip0 = 0; g.qwords = QR; g.twords = TR; g.fwd = 0;
  if (w.pass == 0) {
    g.ncol = (w.algo == ALGO_LOCAL) ? QR * 8 : qlen;     // pad columns cannot reach the global / semi-global result
    g.nstrips = TR;
  } else if (w.algo == ALGO_LOCAL) {                     // local_kernel_template.h:441-451
    g.fwd = a.score[pair];
    if (g.fwd <= 0) { g.ncol = 0; g.nstrips = 0; return g; }
    g.qwords = min((a.q_end[pair] >> 3) + 1, QR); g.twords = min((a.t_end[pair] >> 3) + 1, TR);
    g.ncol = g.qwords * 8; g.nstrips = g.twords;
  } else {                                               // semiglobal_kernel_template.h:273 (unsigned arithmetic)
    g.fwd = a.score[pair];
    const uint32_t ua = (uint32_t)TR - (uint32_t)((a.t_end[pair] >> 3) + 1);
    const uint32_t gend = ua > 0 ? ua - 1 : ua;
    g.ncol = qlen;
    g.strip0 = (int)gend;
    g.nstrips = (gend < (uint32_t)TR) ? TR - (int)gend : 0;
  }
  if (g.ncol <= 0) g.nstrips = 0;
  return g;
}

// query / target word of the sequence as this pass sees it
__device__ __forceinline__ uint32_t wave_qword(const WaveArgs& w, const WaveGeom& g, const uint32_t* qw, int qlen, int j) {
  if (w.pass == 0) return qw[j];
  if (w.algo == ALGO_LOCAL) return nibrev(qw[g.qwords - 1 - j]);
  return rev_true_word(qw, qlen, j);
}
__device__ __forceinline__ uint32_t wave_tword(const WaveArgs& w, const WaveGeom& g, const uint32_t* tw, int tlen, int i /*pass-local strip*/) {
  if (w.pass == 0) return tw[i];
  if (w.algo == ALGO_LOCAL) return nibrev(tw[g.twords - 1 - i]);
  return rev_true_word(tw, tlen, g.strip0 + i);
}

__device__ __forceinline__ int score_general(const Sc& sc, uint32_t qc, uint32_t tc) {
  if (qc == sc.nval || tc == sc.nval) return sc.nscore;
  return (qc == tc) ? sc.match : sc.nmis;
}

// last finished row-block of a pair: fold the per-strip candidates into the result arrays
template <int ALGO>
__device__ __forceinline__ void wave_finalize(const WaveArgs& w, uint32_t pair, const WaveGeom& g, int lane) {
  const DpArgs& a = w.a;
  const int qlen = (int)a.q_lens[pair], tlen = (int)a.t_lens[pair];
  const int2* cand = w.cand + w.soff[pair];
  if (ALGO == ALGO_GLOBAL) return;                        // the capturing lane wrote the score itself
  if (ALGO == ALGO_LOCAL) {
    // first cell in (strip, column, row) order that attains the maximum, strict '>' (:28-29, 412)
    int best = 0, bstrip = 0x7FFFFFFF, bpos = 0;
    for (int i = lane; i < g.nstrips; i += 32) {
      const int2 c = __ldcg(&cand[i]);
      if (c.x > best) { best = c.x; bstrip = i; bpos = c.y; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const int ob = __shfl_xor_sync(0xFFFFFFFFu, best, off), os = __shfl_xor_sync(0xFFFFFFFFu, bstrip, off), op = __shfl_xor_sync(0xFFFFFFFFu, bpos, off);
      if (ob > best || (ob == best && os < bstrip)) { best = ob; bstrip = os; bpos = op; }
    }
    if (lane == 0) {
      const int col = bpos >> 3, row = bpos & 7;
      if (w.pass == 0) {
        a.score[pair] = best;
        a.q_end[pair] = best > 0 ? col : 0;
        a.t_end[pair] = best > 0 ? bstrip * 8 + row : 0;
      } else {
        int sx = 0, sy = 0;
        if (best > 0) { sx = g.ncol - 1 - col; sy = (8 * g.twords - 1 - 8 * bstrip) + row; }     // Q3: mirrored inside the word
        a.q_start[pair] = sx; a.t_start[pair] = sy;
      }
    }
  } else {
    // semi-global, TAIL = TARGET: rows top-down, strict '>' (:160-178); the start pass stops after the first strip that
    // reaches the forward score (:301).  Sequential in strip order; one lane, TR candidates.
    if (lane == 0) {
      if (w.pass == 0) {
        int best = kMinusInf, y = qlen;                                                           // :63-64
        for (int i = 0; i < g.nstrips; i++) { const int2 c = __ldcg(&cand[i]); if (c.x > best) { best = c.x; y = c.y; } }
        a.score[pair] = best; a.t_end[pair] = y; a.q_end[pair] = tlen;                            // Q8
      } else {
        int best = kMinusInf, y = 0;                                                              // :275-276
        for (int i = 0; i < g.nstrips && best < g.fwd; i++) { const int2 c = __ldcg(&cand[i]); if (c.x > best) { best = c.x; y = c.y; } }
        a.t_start[pair] = (tlen - 1) - y;                                                         // :380-381
        a.q_start[pair] = (qlen - 1) - tlen;
      }
    }
  }
}

#ifndef WAVE16_STEP_UNROLL
#define WAVE16_STEP_UNROLL 1
#endif
#define WAVE_PRAGMA_(x) _Pragma(#x)
#define WAVE_UNROLL_(n) WAVE_PRAGMA_(unroll n)
#define WAVE16_UNROLL_PRAGMA WAVE_UNROLL_(WAVE16_STEP_UNROLL)
#ifndef WAVE16_SEMI
#define WAVE16_SEMI 1
#endif
constexpr int kWaveSteadyUnroll = 4;
template <int SR>
struct WaveLane {
  uint32_t tword[SR];           // this lane's packed target words
  uint32_t tbl[8 * SR];         // per row: score against query A, C, G, T (one byte each)
  int p[8 * SR], f[8 * SR], gl[8 * SR];
  int best[SR], bpos[SR];
  int h_bot, e_bot; uint32_t sel_cur;
};

// up to 32 consecutive steps [k0, kend).  STEADY: every lane is inside the matrix for the whole chunk and the scanned /
// captured column (ncol - 1) is not touched, so the body carries no activity or capture tests.  EXOTIC: some column of
// this or the previous chunk holds a query code outside the PRMT table (IUPAC letters; N under the GLOBAL rule).
template <int ALGO, int SR, bool STEADY, bool EXOTIC>
__device__ __forceinline__ void wave_steps(WaveLane<SR>& L, const Sc& sc, uint32_t ntab, const uint2* sin, const uint32_t* ssel, uint2* sout,
                                           int k0, int kend, int ncol, int lane, int qlen, int tlen, int row0, int lstrip, int32_t* score_out) {
  constexpr int ROWS = 8 * SR;
  constexpr bool SEMI = (ALGO == ALGO_SEMI_GLOBAL), LOCALA = (ALGO == ALGO_LOCAL);
  // SEMI and GLOBAL run in a BIASED frame (see dp_s16.cuh nc_sweep4): the value of (lane row r, column c) carries + (c + r) * ext,
  // so a gap extension is free, opening costs `o`, a diagonal step costs + 2 ext (folded into the tables by the caller) and
  // the bottom row leaves the lane with - ROWS * ext.  The reference's finite -infinity (-32768, Q7) is biased like any
  // other value, so its decay by `ext` per step is reproduced exactly.  LOCAL (0 clamp) stays unbiased.
  const int oe = sc.oe, ext = sc.ext, o = sc.o, bias = LOCALA ? 0 : ext;
#pragma unroll (STEADY ? kWaveSteadyUnroll : 1)   // steady chunks are 32 steps: two steps per trip turn the p[r] <- up rotation into renaming
  for (int k = k0; k < kend; k++) {
    // ---- (H, E) of the row above come from the lane above, lane 0 takes them from the chunk ring; every lane reads its
    // column's selector from the 64-column selector ring (address known a step ahead, off the critical path) ----
#ifdef WAVE_SEL_SHFL
    uint32_t sel = __shfl_up_sync(0xFFFFFFFFu, L.sel_cur, 1);
    int h_in = __shfl_up_sync(0xFFFFFFFFu, L.h_bot, 1);
    int e_in = __shfl_up_sync(0xFFFFFFFFu, L.e_bot, 1);
    if (lane == 0) { const uint2 top = sin[k & 31]; h_in = (int)top.x; e_in = (int)top.y; sel = ssel[k & 63]; }
#else
    const uint2 top = sin[k & 31];
    const uint32_t sel = ssel[(k - lane) & 63];
    int h_in = __shfl_up_sync(0xFFFFFFFFu, L.h_bot, 1);
    int e_in = __shfl_up_sync(0xFFFFFFFFu, L.e_bot, 1);
    if (lane == 0) { h_in = (int)top.x; e_in = (int)top.y; }
#endif
    const int c = k - lane;
    if (STEADY || (c >= 0 && c < ncol)) {
#ifdef WAVE_SEL_SHFL
      L.sel_cur = sel;
#endif
      int sv[ROWS];
      if (EXOTIC && (sel & kWaveExotic)) {
        const uint32_t qc = (sel >> 20) & 15u;
#pragma unroll
        for (int s = 0; s < SR; s++)
#pragma unroll
          for (int r = 0; r < 8; r++) sv[8 * s + r] = score_general(sc, qc, nib(L.tword[s], r)) + 2 * bias;
      } else {
#pragma unroll
        for (int r = 0; r < ROWS; r++) sv[r] = (int)prmt(L.tbl[r], ntab, sel);
      }
      int up = h_in, e = e_in;
      int hh[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        const int d = L.p[r] + sv[r];
        int h;
        if (SEMI) {
          L.f[r] = __viaddmax_s32(L.gl[r], -o, L.f[r]);             // F(t,q) = max(H(t,q-1) - oe, F(t,q-1) - e), biased frame
          const int m = max(d, L.f[r]);
          h = max(m, e);
          e = __viaddmax_s32(m, -o, e);                             // E(t+1,q) = max(max(D,F) - oe, E - e)  (oe >= e), biased frame
          L.gl[r] = h;
        } else if (LOCALA) {
          h = __vimax3_s32_relu(d, L.f[r], e);
          const int dm = d - oe;
          L.f[r] = __viaddmax_s32(L.f[r], -ext, dm);
          e = __viaddmax_s32(e, -ext, dm);
        } else {
          h = __vimax3_s32(d, L.f[r], e);
          L.f[r] = __viaddmax_s32(d, -o, L.f[r]);                   // F(t,q+1) = max(F - e, D - oe), biased frame
          e = __viaddmax_s32(d, -o, e);                             // E(t+1,q) = max(E - e, D - oe), biased frame
        }
        hh[r] = h; L.p[r] = up; up = h;
      }
      if (!LOCALA) { up -= ROWS * bias; e -= ROWS * bias; }       // into the frame of the lane below (row -1 / row 0)
      L.h_bot = up; L.e_bot = e;
      if (LOCALA) {
#pragma unroll
        for (int s = 0; s < SR; s++) {
          const int cm = max(__vimax3_s32(hh[8 * s], hh[8 * s + 1], hh[8 * s + 2]),
                             __vimax3_s32(__vimax3_s32(hh[8 * s + 3], hh[8 * s + 4], hh[8 * s + 5]), hh[8 * s + 6], hh[8 * s + 7]));
          if (cm > L.best[s]) {
            int rr = 7;
#pragma unroll
            for (int r = 6; r >= 0; r--) rr = (hh[8 * s + r] == cm) ? r : rr;
            L.best[s] = cm; L.bpos[s] = (c << 3) | rr;
          }
        }
      } else if (!STEADY && c == qlen - 1) {
        if (SEMI) {
          // TAIL == TARGET: last true query column, rows inside the target, strict '>' top-down (:160-178)
#pragma unroll
          for (int s = 0; s < SR; s++)
#pragma unroll
            for (int r = 0; r < 8; r++) {
              const int t = row0 + 8 * s + r, hv = hh[8 * s + r] - (c + 8 * s + r) * bias;
              if (t < tlen && hv > L.best[s]) { L.best[s] = hv; L.bpos[s] = t; }
            }
        } else {
          const int t_last = tlen - 1;
          if ((unsigned)((t_last >> 3) - lstrip) < (unsigned)SR) {
            int v = hh[0] - c * bias;
#pragma unroll
            for (int r = 1; r < ROWS; r++) v = ((t_last - 8 * lstrip) == r) ? hh[r] - (c + r) * bias : v;
            *score_out = v;                                         // H(ref_len-1, read_len-1)  (global.h:285-299)
          }
        }
      }
      if (lane == 31) sout[k & 31] = make_uint2((uint32_t)up, (uint32_t)e);
    }
  }
}

template <int ALGO, int SR>
__global__ void __launch_bounds__(kWaveWarps * 32) k_wave(const WaveArgs w) {
  constexpr bool SEMI = (ALGO == ALGO_SEMI_GLOBAL), LOCALA = (ALGO == ALGO_LOCAL), GLOB = (ALGO == ALGO_GLOBAL);
  __shared__ uint2 s_in[kWaveWarps][32];     // top boundary (H, E) of the current 32-column chunk
  __shared__ uint32_t s_sel[kWaveWarps][64]; // PRMT selectors of the last two chunks' columns
  __shared__ uint2 s_out[kWaveWarps][32];    // bottom boundary produced during the current chunk (slot = step & 31)
  const DpArgs& a = w.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint2* sin = s_in[wib]; uint32_t* ssel = s_sel[wib]; uint2* sout = s_out[wib];
  const Sc sc = make_sc(a, GLOB);
  const int oe = sc.oe, ext = sc.ext;
  const bool top_zero = SEMI && (a.head == SKIP_BOTH);
  const uint32_t total = (uint32_t)a.n * w.maxblk;
  const uint32_t ntab = (uint32_t)(sc.nscore + (LOCALA ? 0 : 2 * ext)) & 0xFFu;   // query N under the LOCAL rule: same score in every row

  for (;;) {
    uint32_t T = 0;
    if (lane == 0) T = atomicAdd(w.ticket, 1u);
    T = __shfl_sync(0xFFFFFFFFu, T, 0);
    if (T >= total) break;
    const uint32_t blk = T / (uint32_t)a.n, pair = T - blk * (uint32_t)a.n;   // block-major tickets (see header)
    const WaveGeom g = wave_geometry(w, pair);
    const int nblk = (g.nstrips + 32 * SR - 1) / (32 * SR);
    if ((int)blk >= nblk) {
      if (blk == 0) wave_finalize<ALGO>(w, pair, g, lane);      // empty problem: defaults
      continue;
    }
    const int qlen = (int)a.q_lens[pair], tlen = (int)a.t_lens[pair];
    const uint32_t* qw = a.pq + (a.q_offs[pair] >> 3);
    const uint32_t* tw = a.pt + (a.t_offs[pair] >> 3);
    const int ncol = g.ncol;
    const int lstrip = ((int)blk * 32 + lane) * SR;             // this lane's first strip, pass-local numbering
    const int row0 = 8 * (g.strip0 + lstrip);                   // its first row in the pass' sequence numbering
    const bool last_blk = ((int)blk == nblk - 1);

    // ---- per-row substitution tables and left-boundary state ----
    WaveLane<SR> L;
#pragma unroll
    for (int s = 0; s < SR; s++) {
      L.tword[s] = (lstrip + s < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip + s) : 0u;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const uint32_t tc = nib(L.tword[s], r);
        const int b2e = LOCALA ? 0 : 2 * ext;                    // biased frame: a diagonal step carries + 2 ext
        const uint32_t b0 = (uint32_t)(score_general(sc, 1u, tc) + b2e) & 0xFFu, b1 = (uint32_t)(score_general(sc, 3u, tc) + b2e) & 0xFFu;
        const uint32_t b2 = (uint32_t)(score_general(sc, 7u, tc) + b2e) & 0xFFu, b3 = (uint32_t)(score_general(sc, 4u, tc) + b2e) & 0xFFu;
        L.tbl[8 * s + r] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        const int t = row0 + 8 * s + r;
        const int lr = 8 * s + r;                                 // row inside the lane: frame of (lr, c) is (c + lr) * ext
        if (LOCALA) { L.p[lr] = 0; L.f[lr] = 0; L.gl[lr] = -oe; }
        else if (GLOB) { L.p[lr] = ((t == 0) ? 0 : -(sc.o + ext * t)) + (lr - 2) * ext; L.f[lr] = kMinusInf + lr * ext; L.gl[lr] = 0; }   // global.h:63-71
        else { L.p[lr] = (lr - 2) * ext; L.f[lr] = kMinusInf + lr * ext; L.gl[lr] = (lr - 1) * ext; }          // free left column (:114-121): H(t,-1) = 0
      }
      L.best[s] = LOCALA ? 0 : kMinusInf; L.bpos[s] = 0;
    }
    L.h_bot = 0; L.e_bot = 0; L.sel_cur = 0;

    const uint4* bnd_in = w.bnd + (uint64_t)((blk + 1) & 1) * w.bnd_stride + w.qoff[pair];
    uint4* bnd_out = w.bnd + (uint64_t)(blk & 1) * w.bnd_stride + w.qoff[pair];
    const uint32_t tag_in = w.tag_base + blk - 1, tag_out = w.tag_base + blk;
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (blk != 0 && lane < ncol) pre = ld_bnd(bnd_in + lane);

    const int nsteps = ncol + 31;
    bool exo_prev = false, exo_cur = false;
    for (int k0 = 0; k0 < nsteps; k0 += 32) {
      // ---- chunk prologue: top boundary and selectors of columns [k0, k0 + 32) into the ring ----
      if (k0 < ncol) {
        const int c = k0 + lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (blk == 0) {
          if (c < ncol) {
            if (GLOB) { v.x = (uint32_t)((c == 0 ? 0 : -(sc.o + ext * c)) + (c - 1) * ext); v.y = (uint32_t)(kMinusInf + c * ext); }   // global.h:57-60, biased
            else if (SEMI) {
              // semi-global hands E over as the first row's own E(0,c) = max(H(-1,c) - oe, E(-1,c) - e)
              const int htop = top_zero ? 0 : (c == 0 ? 0 : -(sc.o + ext * c));                                   // :87-99
              const int etop = top_zero ? 0 : kMinusInf;
              v.x = (uint32_t)(htop + (c - 1) * ext); v.y = (uint32_t)(max(htop - oe, etop - ext) + c * ext);      // biased frame
            }
          }
        } else {
          // the producer's entries validate themselves: (H, tag) (E, tag); re-read until the whole chunk is there
          bool ok = (c >= ncol) || (pre.y == tag_in && pre.w == tag_in);
          while (!__all_sync(0xFFFFFFFFu, ok)) {
            __nanosleep(20);
            if (!ok) { pre = ld_bnd(bnd_in + c); ok = (pre.y == tag_in && pre.w == tag_in); }
          }
          v.x = pre.x; v.y = pre.z;
          if (c + 32 < ncol) pre = ld_bnd(bnd_in + c + 32);      // prefetch the next chunk a whole chunk ahead
        }
        uint32_t selv = 0x8880u;
        if (c < ncol) {
          const uint32_t qword = wave_qword(w, g, qw, qlen, c >> 3);
          const uint32_t qc = nib(qword, c & 7);
          const uint32_t idx = base_idx(qc);
          if (idx < 4) selv = idx * 0x1111u + 0x8880u;
          else if (!GLOB && qc == sc.nval) selv = 0xCCC4u;      // PRMT byte 4 = ntab
          else selv = kWaveExotic | (qc << 20);
        }
        exo_prev = exo_cur;
        exo_cur = __any_sync(0xFFFFFFFFu, (selv & kWaveExotic) != 0);
        __syncwarp();
        sin[lane] = make_uint2(v.x, v.y);
        ssel[c & 63] = selv;
        __syncwarp();
      } else { exo_prev = exo_cur; exo_cur = false; }
      const int kend = min(k0 + 32, nsteps);
      int32_t* sp = a.score + pair;
      if (exo_prev || exo_cur) wave_steps<ALGO, SR, false, true>(L, sc, ntab, sin, ssel, sout, k0, kend, ncol, lane, qlen, tlen, row0, lstrip, sp);
      else if (k0 >= 32 && k0 + 32 < ncol) wave_steps<ALGO, SR, true, false>(L, sc, ntab, sin, ssel, sout, k0, kend, ncol, lane, qlen, tlen, row0, lstrip, sp);
      else wave_steps<ALGO, SR, false, false>(L, sc, ntab, sin, ssel, sout, k0, kend, ncol, lane, qlen, tlen, row0, lstrip, sp);
      // ---- lane 31 finished columns k0-31 .. kend-32 during this chunk: hand them to the next row-block ----
      if (!last_blk) {
        __syncwarp();
        const int cb = k0 - 31 + lane;
        if (cb >= 0 && cb < ncol && k0 + lane < kend) { const uint2 o = sout[lane]; st_bnd(bnd_out + cb, make_uint4(o.x, tag_out, o.y, tag_out)); }
        __syncwarp();
      }
    }
    // ---- per-strip candidates, then the last block of the pair folds them ----
    if (!GLOB) {
      int2* cand = w.cand + w.soff[pair];
#pragma unroll
      for (int s = 0; s < SR; s++)
        if (lstrip + s < g.nstrips) {
          int bp = L.bpos[s];
          if (SEMI && L.best[s] == kMinusInf) bp = 0;
          __stcg(&cand[lstrip + s], make_int2(L.best[s], bp));
        }
      __syncwarp();
      uint32_t old = 0;
      if (lane == 0) { __threadfence(); old = atomicAdd(w.done + pair, 1u); }
      old = __shfl_sync(0xFFFFFFFFu, old, 0);
      if ((int)old == nblk - 1) { __threadfence(); wave_finalize<ALGO>(w, pair, g, lane); }
    }
  }
}

// =====================================================================================================
// k_wave16: the same row-block pipeline with s16x2 lanes for LOCAL alignment (long reads whose scores fit 16 bits).
// A lane owns a DOUBLE-strip: strip 2m in the low 16-bit halves of its registers, strip 2m+1 in the high halves, skewed by
// one query column as in the inter-task fast path (low half on column k - 2*lane, high half one column behind), so the low
// strip's bottom row reaches the high strip through registers and the high strip's bottom row reaches the next lane's low
// strip by one packed __shfl_up_sync.  A row-block is 32 lanes x 16 rows = 512 target rows; per 16 cells a lane issues the
// 8 packed recurrences of dp_s16.cuh (PRMT, 2 x VIADD.16x2, VIMNMX3.S16x2.RELU, 2 x VIADDMNMX.S16x2) -- half the
// instructions per cell of the int32 wavefront kernel.  Halves that are before column 0 or past the last column keep
// computing on benign selectors (score 0 / -1): with the 0 clamp that leaves their state equivalent to the initial one, so
// the step body has no activity tests at all; only the (rare, divergent) record-an-improvement branch checks the column.
struct Wave16Lane {
  uint32_t twA, twB;              // packed target words of the low / high strip
  uint32_t rowA[8], rowB[8];      // per-row substitution tables (dp_s16.cuh row_table layout)
  uint32_t p[8], f[8], hh[8];     // hh: H of the last computed step (its arg-max is taken one step later)
  uint32_t out_h, out_e;          // bottom (H, E) of both strips after the previous step
  uint32_t prevsh;                // selector bits of the high half (the low half's previous column)
  uint32_t vprev;                 // ring word of the previous step (exotic path)
  uint32_t bestpk;                // running per-half maxima, packed (fast improvement test)
  int bestL, bestH, bposL, bposH;
};
// traceback variant: the direction words (8 nibbles = 8 rows of one column) of the last four columns of both halves
struct Wave16Tb { uint32_t dL[4], dH[4]; };

// record an improvement of the running maxima: hh = the H values of step k (low half on column k - 2*lane, high half one behind)
__device__ __forceinline__ void wave16_track(Wave16Lane& L, int k, int ncol, int lane) {
  uint32_t cm = __vimax3_s16x2(L.hh[0], L.hh[1], L.hh[2]);
  cm = __vimax3_s16x2(cm, L.hh[3], L.hh[4]);
  cm = __vimax3_s16x2(cm, L.hh[5], L.hh[6]);
  cm = __vmaxs2(cm, L.hh[7]);
  const uint32_t nb = __vmaxs2(L.bestpk, cm);
  if (nb != L.bestpk) {                              // rare and divergent: resolve row and column, check the column is real
    L.bestpk = nb;
    const int cL = k - 2 * lane, cH = cL - 1;
    const int mL = (int)(short)(cm & 0xFFFFu), mH = (int)(short)(cm >> 16);
    if (mL > L.bestL && cL >= 0 && cL < ncol) {
      int rr = 7;
#pragma unroll
      for (int r = 6; r >= 0; r--) rr = ((int)(short)(L.hh[r] & 0xFFFFu) == mL) ? r : rr;
      L.bestL = mL; L.bposL = (cL << 3) | rr;
    }
    if (mH > L.bestH && cH >= 0 && cH < ncol) {
      int rr = 7;
#pragma unroll
      for (int r = 6; r >= 0; r--) rr = ((int)(short)(L.hh[r] >> 16) == mH) ? r : rr;
      L.bestH = mH; L.bposH = (cH << 3) | rr;
    }
  }
}

template <bool EXOTIC, bool TB = false>
__device__ __forceinline__ void wave16_steps(Wave16Lane& L, const Sc& sc, const uint2* sin, const uint32_t* ssel, uint2* sout,
                                             int k0, int kend, int ncol, int lane, uint32_t neg_oe2, uint32_t neg_ext2,
                                             Wave16Tb* T = nullptr, uint4* __restrict__ tbL = nullptr, uint4* __restrict__ tbH = nullptr, uint32_t mm = 0, uint32_t sixteen = 0) {
WAVE16_UNROLL_PRAGMA
  for (int k = k0; k < kend; k++) {
    const uint2 top = sin[k & 31];
    const uint32_t v = ssel[(k - 2 * lane) & 127];
    uint32_t xh = __shfl_up_sync(0xFFFFFFFFu, L.out_h, 1);
    uint32_t xe = __shfl_up_sync(0xFFFFFFFFu, L.out_e, 1);
    // arg-max of the PREVIOUS step while the shuffles and ring loads are in flight (software-pipelined by one step: the
    // data-dependent branch no longer sits between a step's last row and the next step's shuffle)
    wave16_track(L, k - 1, ncol, lane);
    if (lane == 0) { xh = top.x << 16; xe = top.y << 16; }
    uint32_t up = prmt(xh, L.out_h, 0x5432);       // low half <- lane above, high strip; high half <- own low strip, previous step
    uint32_t e = prmt(xe, L.out_e, 0x5432);
    const uint32_t sel = (v & 0xFFu) | L.prevsh;
    L.prevsh = (v & 0xFFu) * 256u + 0x4400u;
    // the diagonal input of row r is the previous step's H of row r - 1, which L.hh keeps for the arg-max anyway: no second
    // rotating array (8 register moves per step); row 0 takes the previous step's incoming top H from L.p[0]
    uint32_t dg[8];
    dg[0] = L.p[0];
#pragma unroll
    for (int r = 1; r < 8; r++) dg[r] = L.hh[r - 1];
    L.p[0] = up;
    uint32_t wL = 0, wH = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t s2;
      if (EXOTIC && ((v | L.vprev) & kWaveExotic)) {
        const int sl = score_general(sc, (v >> 20) & 15u, nib(L.twA, r)), sh = score_general(sc, (L.vprev >> 20) & 15u, nib(L.twB, r));
        s2 = ((uint32_t)sl & 0xFFFFu) | ((uint32_t)sh << 16);
      } else s2 = prmt(L.rowA[r], L.rowB[r], sel);
      const uint32_t d = __vadd2(dg[r], s2);
      const uint32_t dm = __vadd2(d, neg_oe2);
      const uint32_t h = __vimax3_s16x2_relu(d, L.f[r], e);
      if (TB) tb_nibbles(d, dm, h, L.f[r], e, s2, neg_ext2, 0xFFFFFFFFu, mm, sixteen, wL, wH);
      L.f[r] = __viaddmax_s16x2(L.f[r], neg_ext2, dm);
      e = __viaddmax_s16x2(e, neg_ext2, dm);
      L.hh[r] = h; up = h;
    }
    L.out_h = up; L.out_e = e;
    L.vprev = v;
    if (lane == 31) sout[k & 31] = make_uint2((uint32_t)(int)(short)(up >> 16), (uint32_t)(int)(short)(e >> 16));
    if (TB) {
      // a half completes a tile (4 query columns x 8 rows, one uint4) when its column index reaches 3 mod 4; tiles of one strip
      // are consecutive: tile = strip * 2 QR + (column >> 2) (local_kernel_template.h:257,380), tbL / tbH point at the strips' first tiles
      tb_push(T->dL, wL); tb_push(T->dH, wH);
      const int cL = k - 2 * lane, cH = cL - 1;
      if ((cL & 3) == 3 && cL >= 0 && cL < ncol && tbL) tbL[cL >> 2] = make_uint4(T->dL[0], T->dL[1], T->dL[2], T->dL[3]);
      if ((cH & 3) == 3 && cH >= 0 && cH < ncol && tbH) tbH[cH >> 2] = make_uint4(T->dH[0], T->dH[1], T->dH[2], T->dH[3]);
    }
  }
}

#ifndef WAVE16_MIN_CTAS
#define WAVE16_MIN_CTAS 4
#endif
// TB = true: LOCAL WITH_TB for long pairs -- the same pipeline also emits the reference's 4-bit direction codes
// (local_kernel_template.h:45-60) from the packed lanes (tb_nibbles, dp_common.cuh) in the reference's tile numbering, but into a
// PER-PAIR region of the matrix (DpArgs::tb_off): long pairs differ in size by orders of magnitude, and the reference layout
// (tile t of pair p at tb[t * n + p]) would size every pair for the largest one.  k_get_tb walks the region afterwards.
template <bool TB>
__global__ void __launch_bounds__(kWaveWarps * 32, TB ? 2 : WAVE16_MIN_CTAS) k_wave16_local_t(const WaveArgs w) {
  __shared__ uint2 s_in[kWaveWarps][32];
  __shared__ uint32_t s_sel[kWaveWarps][128];
  __shared__ uint2 s_out[kWaveWarps][32];
  const DpArgs& a = w.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint2* sin = s_in[wib]; uint32_t* ssel = s_sel[wib]; uint2* sout = s_out[wib];
  const Sc sc = make_sc(a, false);
  const uint32_t total = (uint32_t)a.n * w.maxblk;
  const uint32_t neg_oe2 = (((uint32_t)(-sc.oe)) & 0xFFFFu) * 0x00010001u, neg_ext2 = (((uint32_t)(-sc.ext)) & 0xFFFFu) * 0x00010001u;

  for (;;) {
    uint32_t T = 0;
    if (lane == 0) T = atomicAdd(w.ticket, 1u);
    T = __shfl_sync(0xFFFFFFFFu, T, 0);
    if (T >= total) break;
    const uint32_t blk = T / (uint32_t)a.n, pair = T - blk * (uint32_t)a.n;   // block-major tickets (see header)
    const WaveGeom g = wave_geometry(w, pair);
    const int nblk = (g.nstrips + 63) / 64;
    if ((int)blk >= nblk) {
      if (blk == 0) wave_finalize<ALGO_LOCAL>(w, pair, g, lane);
      continue;
    }
    const int qlen = (int)a.q_lens[pair], tlen = (int)a.t_lens[pair];
    const uint32_t* qw = a.pq + (a.q_offs[pair] >> 3);
    const uint32_t* tw = a.pt + (a.t_offs[pair] >> 3);
    const int ncol = g.ncol;
    const int lstrip = ((int)blk * 32 + lane) * 2;
    const bool last_blk = ((int)blk == nblk - 1);

    Wave16Lane L;
    L.twA = (lstrip < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip) : 0u;
    L.twB = (lstrip + 1 < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip + 1) : 0u;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t ta = 0, tb = 0;
      const uint32_t ca = nib(L.twA, r), cb = nib(L.twB, r);
      const uint32_t qcodes[4] = {1u, 3u, 7u, 4u};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        ta |= ((uint32_t)score_general(sc, qcodes[i], ca) & 0xFFu) << (8 * i);
        tb |= ((uint32_t)score_general(sc, qcodes[i], cb) & 0xFFu) << (8 * i);
      }
      L.rowA[r] = ta; L.rowB[r] = tb; L.p[r] = 0; L.f[r] = 0; L.hh[r] = 0;
    }
    L.out_h = 0; L.out_e = 0; L.prevsh = 0xCC00u; L.vprev = 0x88u; L.bestpk = 0; L.bestL = L.bestH = 0; L.bposL = L.bposH = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) ssel[lane + 32 * i] = 0x88u;     // columns before 0 / after the end: score 0 or -1
    __syncwarp();
    Wave16Tb Tb;
    uint4* tbL = nullptr; uint4* tbH = nullptr;
    if (TB) {
#pragma unroll
      for (int i = 0; i < 4; i++) { Tb.dL[i] = 0; Tb.dH[i] = 0; }
      uint4* base = a.tb + a.tb_off[pair];
      const size_t per_strip = (size_t)(ncol >> 2);               // 2 QR tiles per strip
      if (lstrip < g.nstrips) tbL = base + (size_t)lstrip * per_strip;
      if (lstrip + 1 < g.nstrips) tbH = base + (size_t)(lstrip + 1) * per_strip;
    }

    const uint4* bnd_in = w.bnd + (uint64_t)((blk + 1) & 1) * w.bnd_stride + w.qoff[pair];
    uint4* bnd_out = w.bnd + (uint64_t)(blk & 1) * w.bnd_stride + w.qoff[pair];
    const uint32_t tag_in = w.tag_base + blk - 1, tag_out = w.tag_base + blk;
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (blk != 0 && lane < ncol) pre = ld_bnd(bnd_in + lane);

    const int nsteps = ncol + 63;
    int exo_hist = 0;                                            // bit i: chunk (current - i) holds an exotic column
    for (int k0 = 0; k0 < nsteps; k0 += 32) {
      exo_hist <<= 1;
      {
        const int c = k0 + lane;
        uint2 v = make_uint2(0u, 0u);
        uint32_t selv = 0x88u;
        if (k0 < ncol) {
          if (blk != 0) {
            bool ok = (c >= ncol) || (pre.y == tag_in && pre.w == tag_in);
            while (!__all_sync(0xFFFFFFFFu, ok)) {
              __nanosleep(20);
              if (!ok) { pre = ld_bnd(bnd_in + c); ok = (pre.y == tag_in && pre.w == tag_in); }
            }
            v.x = pre.x; v.y = pre.z;
            if (c + 32 < ncol) pre = ld_bnd(bnd_in + c + 32);
          }
          if (c < ncol) {
            const uint32_t qword = wave_qword(w, g, qw, qlen, c >> 3);
            const uint32_t qc = nib(qword, c & 7);
            const uint32_t idx = base_idx(qc);
            selv = (idx < 4) ? (0x80u + 0x11u * idx) | (qc << 20) : (0x88u | kWaveExotic | (qc << 20));
          }
        }
        if (__any_sync(0xFFFFFFFFu, (selv & kWaveExotic) != 0)) exo_hist |= 1;
        __syncwarp();
        sin[lane] = v;
        ssel[c & 127] = selv;
        __syncwarp();
      }
      const int kend = min(k0 + 32, nsteps);
      if (exo_hist & 7) wave16_steps<true, TB>(L, sc, sin, ssel, sout, k0, kend, ncol, lane, neg_oe2, neg_ext2, &Tb, tbL, tbH, w.mm, w.sixteen);
      else wave16_steps<false, TB>(L, sc, sin, ssel, sout, k0, kend, ncol, lane, neg_oe2, neg_ext2, &Tb, tbL, tbH, w.mm, w.sixteen);
      // ---- lane 31's high strip finished columns k0-63 .. kend-64 during this chunk ----
      if (!last_blk) {
        __syncwarp();
        const int cb = k0 - 63 + lane;
        if (cb >= 0 && cb < ncol && k0 + lane < kend) { const uint2 o = sout[lane]; st_bnd(bnd_out + cb, make_uint4(o.x, tag_out, o.y, tag_out)); }
        __syncwarp();
      }
    }
    wave16_track(L, nsteps - 1, ncol, lane);                    // the last step's arg-max is still pending
    int2* cand = w.cand + w.soff[pair];
    if (lstrip < g.nstrips) __stcg(&cand[lstrip], make_int2(L.bestL, L.bposL));
    if (lstrip + 1 < g.nstrips) __stcg(&cand[lstrip + 1], make_int2(L.bestH, L.bposH));
    __syncwarp();
    uint32_t old = 0;
    if (lane == 0) { __threadfence(); old = atomicAdd(w.done + pair, 1u); }
    old = __shfl_sync(0xFFFFFFFFu, old, 0);
    if ((int)old == nblk - 1) { __threadfence(); wave_finalize<ALGO_LOCAL>(w, pair, g, lane); }
  }
}


// =====================================================================================================
// k_wave16_global: GLOBAL score (global.h) on the s16x2 row-block pipeline of k_wave16_local.  A 16k x 16k global matrix
// does not fit 16 bits as it stands (H runs from -(o + 16384 e) to +16384 match), but in the BIASED frame of dp_s16.cuh every
// value carries + (column + row-in-strip) * ext, which cancels the linear decay of the gap costs, and a constant shift K
// (recurrences only add and max: a constant commutes with both) centres what is left: wave16_global_range() proves on the
// host, from the batch's lengths and the scoring, that every real value stays inside (floor16, 32767) -- else the batch runs
// on the int32 kernel.  Per cell pair: PRMT, VIADD.16x2, VIMNMX3.S16x2, 2 x VIADDMNMX.S16x2 (5 instructions for two cells;
// the int32 kernel needs 5 per cell).  There is no 0 clamp to neutralise the columns a half computes before its column 0,
// so each half gets its strip-start state (diagonal inputs, F = -infinity) imposed at the step it enters the matrix.
struct Wave16G {
  uint32_t rowA[8], rowB[8];
  uint32_t f[8], hh[8];           // hh: H of the previous step = diagonal input of the row below
  uint32_t p0;                    // incoming top H of the previous step = diagonal input of row 0
  uint32_t out_h, out_e, prevsh, vprev;
  uint32_t twA, twB;
};

template <bool EXOTIC>
__device__ __forceinline__ void wave16g_steps(Wave16G& L, const Sc& sc, const uint2* sin, const uint32_t* ssel, uint2* sout,
                                              int k0, int kend, int lane, uint32_t neg_o2, uint32_t conv8, const uint32_t (&dinit)[8], uint32_t floor2,
                                              int kcapL, int kcapH, int caprow, int unbias_cap, int32_t* score_out) {
WAVE16_UNROLL_PRAGMA
  for (int k = k0; k < kend; k++) {
    const uint2 top = sin[k & 31];
    const uint32_t v = ssel[(k - 2 * lane) & 127];
    uint32_t xh = __shfl_up_sync(0xFFFFFFFFu, L.out_h, 1);
    uint32_t xe = __shfl_up_sync(0xFFFFFFFFu, L.out_e, 1);
    if (lane == 0) { xh = top.x << 16; xe = top.y << 16; }
    uint32_t up = prmt(xh, L.out_h, 0x5432);       // low half <- lane above, high strip; high half <- own low strip, previous step
    uint32_t e = prmt(xe, L.out_e, 0x5432);
    const uint32_t sel = (v & 0xFFu) | L.prevsh;
    L.prevsh = (v & 0xFFu) * 256u + 0x4400u;
    // a half enters the matrix: low half at step 2 * lane (column 0), high half one step later
    const int kin = k - 2 * lane;
    if ((unsigned)kin < 2u) {
      const uint32_t m = kin ? 0x7610u : 0x3254u;  // PRMT selector: replace the high (kin = 1) or the low (kin = 0) half by the second operand's
      L.p0 = prmt(L.p0, dinit[0], m);
#pragma unroll
      for (int r = 0; r < 7; r++) L.hh[r] = prmt(L.hh[r], dinit[r + 1], m);
#pragma unroll
      for (int r = 0; r < 8; r++) L.f[r] = prmt(L.f[r], floor2, m);
    }
    uint32_t dg[8];
    dg[0] = L.p0;
#pragma unroll
    for (int r = 1; r < 8; r++) dg[r] = L.hh[r - 1];
    L.p0 = up;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t s2;
      if (EXOTIC && ((v | L.vprev) & kWaveExotic)) {
        const int sl = score_general(sc, (v >> 20) & 15u, nib(L.twA, r)) + 2 * sc.ext, sh = score_general(sc, (L.vprev >> 20) & 15u, nib(L.twB, r)) + 2 * sc.ext;
        s2 = ((uint32_t)sl & 0xFFFFu) | ((uint32_t)sh << 16);
      } else s2 = prmt(L.rowA[r], L.rowB[r], sel);
      const uint32_t d = __vadd2(dg[r], s2);
      const uint32_t h = __vimax3_s16x2(d, L.f[r], e);
      L.f[r] = __viaddmax_s16x2(d, neg_o2, L.f[r]);
      e = __viaddmax_s16x2(d, neg_o2, e);
      L.hh[r] = h; up = h;
    }
    if (k == kcapL || k == kcapH) {                                                         // H(ref_len-1, read_len-1)  (global.h:285-299)
      uint32_t hv = L.hh[0];
#pragma unroll
      for (int r = 1; r < 8; r++) hv = (caprow == r) ? L.hh[r] : hv;                        // select chain: no dynamic register indexing
      *score_out = ((k == kcapL) ? (int)(short)(hv & 0xFFFFu) : (int)(short)(hv >> 16)) - unbias_cap;
    }
    L.out_h = __vadd2(up, conv8); L.out_e = __vadd2(e, conv8);                              // into the frame of the strip below
    L.vprev = v;
    if (lane == 31) sout[k & 31] = make_uint2((uint32_t)(int)(short)(L.out_h >> 16), (uint32_t)(int)(short)(L.out_e >> 16));
  }
}

__global__ void __launch_bounds__(kWaveWarps * 32, WAVE16_MIN_CTAS) k_wave16_global(const WaveArgs w) {
  __shared__ uint2 s_in[kWaveWarps][32];
  __shared__ uint32_t s_sel[kWaveWarps][128];
  __shared__ uint2 s_out[kWaveWarps][32];
  const DpArgs& a = w.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint2* sin = s_in[wib]; uint32_t* ssel = s_sel[wib]; uint2* sout = s_out[wib];
  const Sc sc = make_sc(a, true);
  const int ext = sc.ext, o = sc.o, K = w.kshift;
  const uint32_t total = (uint32_t)a.n * w.maxblk;
  const uint32_t neg_o2 = (((uint32_t)(-o)) & 0xFFFFu) * 0x00010001u, conv8 = (((uint32_t)(-8 * ext)) & 0xFFFFu) * 0x00010001u;
  const uint32_t floor2 = (((uint32_t)w.floor16) & 0xFFFFu) * 0x00010001u;

  for (;;) {
    uint32_t T = 0;
    if (lane == 0) T = atomicAdd(w.ticket, 1u);
    T = __shfl_sync(0xFFFFFFFFu, T, 0);
    if (T >= total) break;
    const uint32_t blk = T / (uint32_t)a.n, pair = T - blk * (uint32_t)a.n;
    const WaveGeom g = wave_geometry(w, pair);
    const int nblk = (g.nstrips + 63) / 64;
    if ((int)blk >= nblk) continue;
    const int qlen = (int)a.q_lens[pair], tlen = (int)a.t_lens[pair];
    const uint32_t* qw = a.pq + (a.q_offs[pair] >> 3);
    const uint32_t* tw = a.pt + (a.t_offs[pair] >> 3);
    const int ncol = g.ncol;
    const int lstrip = ((int)blk * 32 + lane) * 2;
    const bool last_blk = ((int)blk == nblk - 1);

    Wave16G L;
    L.twA = (lstrip < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip) : 0u;
    L.twB = (lstrip + 1 < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip + 1) : 0u;
    uint32_t dinit[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t ta = 0, tb = 0;
      const uint32_t ca = nib(L.twA, r), cb = nib(L.twB, r);
      const uint32_t qcodes[4] = {1u, 3u, 7u, 4u};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        ta |= ((uint32_t)(score_general(sc, qcodes[i], ca) + 2 * ext) & 0xFFu) << (8 * i);
        tb |= ((uint32_t)(score_general(sc, qcodes[i], cb) + 2 * ext) & 0xFFu) << (8 * i);
      }
      L.rowA[r] = ta; L.rowB[r] = tb; L.f[r] = floor2; L.hh[r] = 0;
      // diagonal input of (row r, column 0) = H(t - 1, -1) = t ? -(o + e t) : 0 (global.h:66-70), in the frame of (r - 1, -1), shifted by K
      const int tA = 8 * lstrip + r, tB = tA + 8, fb = (r - 2) * ext - K;
      dinit[r] = ((uint32_t)(((tA == 0) ? 0 : -(o + ext * tA)) + fb) & 0xFFFFu) | ((uint32_t)(-(o + ext * tB) + fb) << 16);
    }
    L.p0 = 0; L.out_h = 0; L.out_e = 0; L.prevsh = 0xCC00u; L.vprev = 0x88u;
#pragma unroll
    for (int i = 0; i < 4; i++) ssel[lane + 32 * i] = 0x88u;
    __syncwarp();
    // which half of which lane holds row tlen - 1, and at which step it is on column qlen - 1
    const int t_last = tlen - 1, sl = t_last >> 3, caprow = t_last & 7;
    const int kcapL = (sl == lstrip) ? (qlen - 1) + 2 * lane : -1, kcapH = (sl == lstrip + 1) ? (qlen - 1) + 2 * lane + 1 : -1;
    const int unbias_cap = (qlen - 1 + caprow) * ext - K;

    const uint4* bnd_in = w.bnd + (uint64_t)((blk + 1) & 1) * w.bnd_stride + w.qoff[pair];
    uint4* bnd_out = w.bnd + (uint64_t)(blk & 1) * w.bnd_stride + w.qoff[pair];
    const uint32_t tag_in = w.tag_base + blk - 1, tag_out = w.tag_base + blk;
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (blk != 0 && lane < ncol) pre = ld_bnd(bnd_in + lane);

    const int nsteps = ncol + 63;
    int exo_hist = 0;
    for (int k0 = 0; k0 < nsteps; k0 += 32) {
      exo_hist <<= 1;
      {
        const int c = k0 + lane;
        uint2 v = make_uint2(0u, 0u);
        uint32_t selv = 0x88u;
        if (k0 < ncol) {
          if (blk != 0) {
            bool ok = (c >= ncol) || (pre.y == tag_in && pre.w == tag_in);
            while (!__all_sync(0xFFFFFFFFu, ok)) {
              __nanosleep(20);
              if (!ok) { pre = ld_bnd(bnd_in + c); ok = (pre.y == tag_in && pre.w == tag_in); }
            }
            v.x = pre.x; v.y = pre.z;
            if (c + 32 < ncol) pre = ld_bnd(bnd_in + c + 32);
          } else if (c < ncol) {
            // top boundary (global.h:57-60): H(-1,c) = c ? -(o + e c) : 0 in the frame of (row -1, column c); E(0,c) = -infinity
            v.x = (uint32_t)(((c == 0) ? 0 : -(o + ext * c)) + (c - 1) * ext - K); v.y = (uint32_t)w.floor16;
          }
          if (c < ncol) {
            const uint32_t qword = wave_qword(w, g, qw, qlen, c >> 3);
            const uint32_t qc = nib(qword, c & 7);
            const uint32_t idx = base_idx(qc);
            selv = (idx < 4) ? (0x80u + 0x11u * idx) | (qc << 20) : (0x88u | kWaveExotic | (qc << 20));
          }
        }
        if (__any_sync(0xFFFFFFFFu, (selv & kWaveExotic) != 0)) exo_hist |= 1;
        __syncwarp();
        sin[lane] = v;
        ssel[c & 127] = selv;
        __syncwarp();
      }
      const int kend = min(k0 + 32, nsteps);
      int32_t* sp = a.score + pair;
      if (exo_hist & 7) wave16g_steps<true>(L, sc, sin, ssel, sout, k0, kend, lane, neg_o2, conv8, dinit, floor2, kcapL, kcapH, caprow, unbias_cap, sp);
      else wave16g_steps<false>(L, sc, sin, ssel, sout, k0, kend, lane, neg_o2, conv8, dinit, floor2, kcapL, kcapH, caprow, unbias_cap, sp);
      if (!last_blk) {
        __syncwarp();
        const int cb = k0 - 63 + lane;
        if (cb >= 0 && cb < ncol && k0 + lane < kend) { const uint2 ov = sout[lane]; st_bnd(bnd_out + cb, make_uint4(ov.x, tag_out, ov.y, tag_out)); }
        __syncwarp();
      }
    }
  }
}

// =====================================================================================================
// k_wave16_semi: SEMI_GLOBAL (TAIL = TARGET, HEAD in {TARGET, BOTH}: semiglobal_kernel_template.h:39-388) on the same s16x2
// row-block pipeline, in the biased and shifted frame of k_wave16_global.  Recurrence of dp_s16.cuh nc_step (H-based opening
// without the H -> E -> H chain; E parked as E + open): per cell pair PRMT, VIADDMNMX (F from the previous column's H),
// VIMNMX3 (H), VIMNMX3 (E + open) + two VIADD.16x2.  Free left column: every half enters the matrix with H(r, -1) = 0.  The last
// true query column is scanned per strip (rows inside the target, strict '>' top-down, :160-178); the last finishing block folds
// the per-strip candidates (wave_finalize: the start pass stops after the first strip that reaches the forward score, :301).
struct Wave16S {
  uint32_t rowA[8], rowB[8];
  uint32_t f[8], hh[8];           // hh: H of the previous step = F input of the same row and diagonal input of the row below
  uint32_t p0;                    // incoming top H of the previous step = diagonal input of row 0
  uint32_t out_h, out_e, prevsh, vprev;
  uint32_t twA, twB;
  int bestL, bestH, yL, yH;       // per-strip candidates: best H(row, qlen - 1) and its row (sequence numbering of the pass)
};

template <bool EXOTIC>
__device__ __forceinline__ void wave16s_steps(Wave16S& L, const Sc& sc, const uint2* sin, const uint32_t* ssel, uint2* sout,
                                              int k0, int kend, int lane, uint32_t neg_o2, uint32_t conv8, const uint32_t (&dinit)[9], uint32_t floor2,
                                              int kscanL, int kscanH, int rowbaseL, int rowbaseH, int tlen, int unbias0) {
WAVE16_UNROLL_PRAGMA
  for (int k = k0; k < kend; k++) {
    const uint2 top = sin[k & 31];
    const uint32_t v = ssel[(k - 2 * lane) & 127];
    uint32_t xh = __shfl_up_sync(0xFFFFFFFFu, L.out_h, 1);
    uint32_t xe = __shfl_up_sync(0xFFFFFFFFu, L.out_e, 1);
    if (lane == 0) { xh = top.x << 16; xe = top.y << 16; }
    uint32_t up = prmt(xh, L.out_h, 0x5432);       // low half <- lane above, high strip; high half <- own low strip, previous step
    uint32_t e = prmt(xe, L.out_e, 0x5432);
    const uint32_t sel = (v & 0xFFu) | L.prevsh;
    L.prevsh = (v & 0xFFu) * 256u + 0x4400u;
    // a half enters the matrix: low half at step 2 * lane (column 0), high half one step later: impose the free left column
    const int kin = k - 2 * lane;
    if ((unsigned)kin < 2u) {
      const uint32_t m = kin ? 0x7610u : 0x3254u;
      L.p0 = prmt(L.p0, dinit[0], m);
#pragma unroll
      for (int r = 0; r < 8; r++) L.hh[r] = prmt(L.hh[r], dinit[r + 1], m);    // H(r, -1) = 0 in the frame of (r, -1): F input of row r, diagonal input of row r + 1
#pragma unroll
      for (int r = 0; r < 8; r++) L.f[r] = prmt(L.f[r], floor2, m);
    }
    uint32_t dg[8];
    dg[0] = L.p0;
#pragma unroll
    for (int r = 1; r < 8; r++) dg[r] = L.hh[r - 1];
    L.p0 = up;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t s2;
      if (EXOTIC && ((v | L.vprev) & kWaveExotic)) {
        const int sl = score_general(sc, (v >> 20) & 15u, nib(L.twA, r)) + 2 * sc.ext, sh = score_general(sc, (L.vprev >> 20) & 15u, nib(L.twB, r)) + 2 * sc.ext;
        s2 = ((uint32_t)sl & 0xFFFFu) | ((uint32_t)sh << 16);
      } else s2 = prmt(L.rowA[r], L.rowB[r], sel);
      const uint32_t d = __vadd2(dg[r], s2);
      L.f[r] = __viaddmax_s16x2(L.hh[r], neg_o2, L.f[r]);              // F(t,q) = max(H(t,q-1) - oe, F(t,q-1) - e), biased frame
      const uint32_t h = __vimax3_s16x2(d, L.f[r], __vadd2(e, neg_o2)); // e holds E + open
      e = __vimax3_s16x2(e, d, L.f[r]);                                // E(t+1,q) + open = max(E + open, max(D, F)), biased frame
      L.hh[r] = h; up = h;
    }
    if (k == kscanL || k == kscanH) {                                  // this half is on the last true query column (:160-178)
      const bool low = (k == kscanL);
      const int rb = low ? rowbaseL : rowbaseH;
      int best = low ? L.bestL : L.bestH, y = low ? L.yL : L.yH;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int hv = (low ? (int)(short)(L.hh[r] & 0xFFFFu) : (int)(short)(L.hh[r] >> 16)) - (unbias0 + r * sc.ext);
        if (rb + r < tlen && hv > best) { best = hv; y = rb + r; }
      }
      if (low) { L.bestL = best; L.yL = y; } else { L.bestH = best; L.yH = y; }
    }
    L.out_h = __vadd2(up, conv8); L.out_e = __vadd2(e, conv8);          // into the frame of the strip below
    L.vprev = v;
    if (lane == 31) sout[k & 31] = make_uint2((uint32_t)(int)(short)(L.out_h >> 16), (uint32_t)(int)(short)(L.out_e >> 16));
  }
}

__global__ void __launch_bounds__(kWaveWarps * 32, WAVE16_MIN_CTAS) k_wave16_semi(const WaveArgs w) {
  __shared__ uint2 s_in[kWaveWarps][32];
  __shared__ uint32_t s_sel[kWaveWarps][128];
  __shared__ uint2 s_out[kWaveWarps][32];
  const DpArgs& a = w.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint2* sin = s_in[wib]; uint32_t* ssel = s_sel[wib]; uint2* sout = s_out[wib];
  const Sc sc = make_sc(a, false);
  const int ext = sc.ext, o = sc.o, K = w.kshift;
  const bool top_zero = (a.head == SKIP_BOTH);
  const uint32_t total = (uint32_t)a.n * w.maxblk;
  const uint32_t neg_o2 = (((uint32_t)(-o)) & 0xFFFFu) * 0x00010001u, conv8 = (((uint32_t)(-8 * ext)) & 0xFFFFu) * 0x00010001u;
  const uint32_t floor2 = (((uint32_t)w.floor16) & 0xFFFFu) * 0x00010001u;

  for (;;) {
    uint32_t T = 0;
    if (lane == 0) T = atomicAdd(w.ticket, 1u);
    T = __shfl_sync(0xFFFFFFFFu, T, 0);
    if (T >= total) break;
    const uint32_t blk = T / (uint32_t)a.n, pair = T - blk * (uint32_t)a.n;
    const WaveGeom g = wave_geometry(w, pair);
    const int nblk = (g.nstrips + 63) / 64;
    if ((int)blk >= nblk) {
      if (blk == 0) wave_finalize<ALGO_SEMI_GLOBAL>(w, pair, g, lane);   // empty problem: defaults
      continue;
    }
    const int qlen = (int)a.q_lens[pair], tlen = (int)a.t_lens[pair];
    const uint32_t* qw = a.pq + (a.q_offs[pair] >> 3);
    const uint32_t* tw = a.pt + (a.t_offs[pair] >> 3);
    const int ncol = g.ncol;
    const int lstrip = ((int)blk * 32 + lane) * 2;
    const bool last_blk = ((int)blk == nblk - 1);

    Wave16S L;
    L.twA = (lstrip < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip) : 0u;
    L.twB = (lstrip + 1 < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip + 1) : 0u;
    uint32_t dinit[9];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t ta = 0, tb = 0;
      const uint32_t ca = nib(L.twA, r), cb = nib(L.twB, r);
      const uint32_t qcodes[4] = {1u, 3u, 7u, 4u};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        ta |= ((uint32_t)(score_general(sc, qcodes[i], ca) + 2 * ext) & 0xFFu) << (8 * i);
        tb |= ((uint32_t)(score_general(sc, qcodes[i], cb) + 2 * ext) & 0xFFu) << (8 * i);
      }
      L.rowA[r] = ta; L.rowB[r] = tb; L.f[r] = floor2; L.hh[r] = 0;
    }
    // free left column (:114-121): dinit[r] = H(r - 1, -1) = 0 in the frame of (r - 1, -1), shifted by K
#pragma unroll
    for (int r = 0; r < 9; r++) dinit[r] = (((uint32_t)((r - 2) * ext - K)) & 0xFFFFu) * 0x00010001u;
    L.p0 = 0; L.out_h = 0; L.out_e = 0; L.prevsh = 0xCC00u; L.vprev = 0x88u;
    L.bestL = L.bestH = kMinusInf; L.yL = L.yH = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) ssel[lane + 32 * i] = 0x88u;
    __syncwarp();
    const int kscanL = (lstrip < g.nstrips) ? (qlen - 1) + 2 * lane : -1, kscanH = (lstrip + 1 < g.nstrips) ? (qlen - 1) + 2 * lane + 1 : -1;
    const int rowbaseL = 8 * (g.strip0 + lstrip), rowbaseH = rowbaseL + 8;
    const int unbias0 = (qlen - 1) * ext - K;

    const uint4* bnd_in = w.bnd + (uint64_t)((blk + 1) & 1) * w.bnd_stride + w.qoff[pair];
    uint4* bnd_out = w.bnd + (uint64_t)(blk & 1) * w.bnd_stride + w.qoff[pair];
    const uint32_t tag_in = w.tag_base + blk - 1, tag_out = w.tag_base + blk;
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (blk != 0 && lane < ncol) pre = ld_bnd(bnd_in + lane);

    const int nsteps = ncol + 63;
    int exo_hist = 0;
    for (int k0 = 0; k0 < nsteps; k0 += 32) {
      exo_hist <<= 1;
      {
        const int c = k0 + lane;
        uint2 v = make_uint2(0u, 0u);
        uint32_t selv = 0x88u;
        if (k0 < ncol) {
          if (blk != 0) {
            bool ok = (c >= ncol) || (pre.y == tag_in && pre.w == tag_in);
            while (!__all_sync(0xFFFFFFFFu, ok)) {
              __nanosleep(20);
              if (!ok) { pre = ld_bnd(bnd_in + c); ok = (pre.y == tag_in && pre.w == tag_in); }
            }
            v.x = pre.x; v.y = pre.z;
            if (c + 32 < ncol) pre = ld_bnd(bnd_in + c + 32);
          } else if (c < ncol) {
            // top boundary (:87-99): H(-1,c) = c ? -(o + e c) : 0 (HEAD = BOTH: 0) in the frame of (row -1, column c); the first row's
            // own E(0,c) = H(-1,c) - oe (HEAD = BOTH: max(0 - oe, 0 - e) = -e), parked as E + open in the frame of (row 0, column c)
            const int htop = top_zero ? 0 : ((c == 0) ? 0 : -(o + ext * c));
            v.x = (uint32_t)(htop + (c - 1) * ext - K);
            v.y = (uint32_t)((top_zero ? -ext : htop - (o + ext)) + c * ext + o - K);
          }
          if (c < ncol) {
            const uint32_t qword = wave_qword(w, g, qw, qlen, c >> 3);
            const uint32_t qc = nib(qword, c & 7);
            const uint32_t idx = base_idx(qc);
            selv = (idx < 4) ? (0x80u + 0x11u * idx) | (qc << 20) : (0x88u | kWaveExotic | (qc << 20));
          }
        }
        if (__any_sync(0xFFFFFFFFu, (selv & kWaveExotic) != 0)) exo_hist |= 1;
        __syncwarp();
        sin[lane] = v;
        ssel[c & 127] = selv;
        __syncwarp();
      }
      const int kend = min(k0 + 32, nsteps);
      if (exo_hist & 7) wave16s_steps<true>(L, sc, sin, ssel, sout, k0, kend, lane, neg_o2, conv8, dinit, floor2, kscanL, kscanH, rowbaseL, rowbaseH, tlen, unbias0);
      else wave16s_steps<false>(L, sc, sin, ssel, sout, k0, kend, lane, neg_o2, conv8, dinit, floor2, kscanL, kscanH, rowbaseL, rowbaseH, tlen, unbias0);
      if (!last_blk) {
        __syncwarp();
        const int cb = k0 - 63 + lane;
        if (cb >= 0 && cb < ncol && k0 + lane < kend) { const uint2 ov = sout[lane]; st_bnd(bnd_out + cb, make_uint4(ov.x, tag_out, ov.y, tag_out)); }
        __syncwarp();
      }
    }
    int2* cand = w.cand + w.soff[pair];
    if (lstrip < g.nstrips) __stcg(&cand[lstrip], make_int2(L.bestL, L.bestL == kMinusInf ? 0 : L.yL));
    if (lstrip + 1 < g.nstrips) __stcg(&cand[lstrip + 1], make_int2(L.bestH, L.bestH == kMinusInf ? 0 : L.yH));
    __syncwarp();
    uint32_t old = 0;
    if (lane == 0) { __threadfence(); old = atomicAdd(w.done + pair, 1u); }
    old = __shfl_sync(0xFFFFFFFFu, old, 0);
    if ((int)old == nblk - 1) { __threadfence(); wave_finalize<ALGO_SEMI_GLOBAL>(w, pair, g, lane); }
  }
}

// Range proof for k_wave16_semi: biased values X + (c + r) ext - K of every real H / E + open / F.  The free left column makes the
// lower end independent of the lengths (every cell can be reached from H(t, -1) = 0 by one gap of c + 1 columns, which the frame
// cancels): lower >= -(3 oe + mismatch + 9 ext + 64); upper: H <= match * min(q, t), bias <= (q8 + 64 + 8) ext, E carries + open.
inline bool wave16_semi_range(const DpArgs& a, uint32_t max_q, uint32_t max_t, int* kshift, int* floor16) {
  if (a.match < 0 || a.mismatch < 0 || a.gap_open < 0 || a.gap_extend < 0 || a.match + 2 * a.gap_extend > 127) return false;
  if (a.n_score < -127 || a.n_score + 2 * a.gap_extend > 127) return false;
  const int64_t q8 = (max_q + 7) & ~7ll, t8 = (max_t + 7) & ~7ll, oe = (int64_t)a.gap_open + a.gap_extend;
  const int64_t ub = (int64_t)a.match * std::min<int64_t>(q8, t8) + (q8 + 72) * a.gap_extend + a.gap_open + 64;
  const int64_t lb = -(3 * oe + (int64_t)a.mismatch + 9 * a.gap_extend + std::max(0, -a.n_score) + 64);
  const int64_t K = (ub + lb) / 2, fl = lb - K - 1000;
  if (ub - K > 32000 || fl < -32700 || K > 32000 || K < -32000) return false;
  *kshift = (int)K; *floor16 = (int)fl;
  return true;
}

// Range proof for k_wave16_global: biased values X + (c + r) ext - K of every real H / E / F, for all pairs of the batch.
//   upper: H <= match * min(q, t), bias <= (q8 + 64 + 8) ext        lower: H, E, F >= -(2 oe + ext * (t8 + 8) + mismatch + 64) (bias >= 0 cancels the query side)
// The shift K centres the interval; the floor sits 1000 below the lower end and must itself fit.
inline bool wave16_global_range(const DpArgs& a, uint32_t max_q, uint32_t max_t, int* kshift, int* floor16) {
  if (a.match < 0 || a.mismatch < 0 || a.gap_open < 0 || a.gap_extend < 0 || a.match + 2 * a.gap_extend > 127) return false;
  if (a.n_is_wild_global && (a.n_score < -127 || a.n_score + 2 * a.gap_extend > 127)) return false;
  const int64_t q8 = (max_q + 7) & ~7ll, t8 = (max_t + 7) & ~7ll, oe = (int64_t)a.gap_open + a.gap_extend;
  const int64_t ub = (int64_t)a.match * std::min<int64_t>(q8, t8) + (q8 + 72) * a.gap_extend + 64;
  const int64_t lb = -(2 * oe + (int64_t)a.gap_extend * (t8 + 8) + a.mismatch + 64);
  const int64_t K = (ub + lb) / 2, fl = lb - K - 1000;
  if (ub - K > 32000 || fl < -32700 || K > 32000 || K < -32000) return false;
  *kshift = (int)K; *floor16 = (int)fl;
  return true;
}

bool wave16_eligible(const DpArgs& a, int algo, uint32_t max_q, uint32_t max_t) {
  if (algo == ALGO_GLOBAL) { int k, f; return wave16_global_range(a, max_q, max_t, &k, &f); }
  if (algo == ALGO_SEMI_GLOBAL) {
    int k, f;
    return WAVE16_SEMI && a.tail == SKIP_TARGET && (a.head == SKIP_TARGET || a.head == SKIP_BOTH) && wave16_semi_range(a, max_q, max_t, &k, &f);
  }
  if (algo != ALGO_LOCAL) return false;
  if ((uint64_t)std::min(max_q, max_t) * (uint64_t)a.match > 32000ull) return false;          // scores stay in 16 bits
  return a.gap_open + a.gap_extend + a.mismatch + std::max(0, -a.n_score) < 8000;
}

// ---- host side (WaveScratch: dp_launch.h) ----------------------------------------------------------------------
bool wave_eligible(const DpArgs& a, int algo, int start, bool second, uint32_t min_q, uint32_t min_t) {
  if (second) return false;
  if (start == START_TB && algo != ALGO_LOCAL) return false;   // traceback: LOCAL on the s16x2 lanes only (the caller checks wave16_eligible)
  if (min_q == 0 || min_t == 0) return false;
  if (a.match < 0 || a.match > 127 || a.mismatch < 0 || a.mismatch > 127) return false;
  if (a.n_score < -127) return false;
  if (a.gap_open < 0 || a.gap_extend < 0) return false;
  if (algo == ALGO_LOCAL) return true;
  if (a.match + 2 * a.gap_extend > 127) return false;          // biased frame: the substitution bytes carry + 2 ext
  if (algo == ALGO_GLOBAL) return start == START_NONE;
  if (algo == ALGO_SEMI_GLOBAL) return a.tail == SKIP_TARGET && (a.head == SKIP_TARGET || a.head == SKIP_BOTH);
  return false;
}

template <int SR>
inline void wave_launch_kernel(int algo, int blocks, const WaveArgs& wa, cudaStream_t str) {
  if (algo == ALGO_LOCAL) k_wave<ALGO_LOCAL, SR><<<blocks, kWaveWarps * 32, 0, str>>>(wa);
  else if (algo == ALGO_GLOBAL) k_wave<ALGO_GLOBAL, SR><<<blocks, kWaveWarps * 32, 0, str>>>(wa);
  else k_wave<ALGO_SEMI_GLOBAL, SR><<<blocks, kWaveWarps * 32, 0, str>>>(wa);
}

// h_qlens / h_tlens: the batch's lengths on the host (the caller's pinned arrays).  Returns the number of kernel
// launches, or -1 when a scratch allocation failed.
int wave_launch(const DpArgs& a, int algo, int start, WaveScratch& ws, const uint32_t* h_qlens, const uint32_t* h_tlens,
                uint32_t max_t, int n_sm, cudaStream_t str, bool* s16_inout, bool force_s16) {
  const size_t n = (size_t)a.n;
  if (ws.h_cap < 3 * n) {
    if (ws.h_off) cudaFreeHost(ws.h_off);
    ws.h_off = nullptr;
    if (cudaHostAlloc(&ws.h_off, 3 * n * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess) return -1;
    ws.h_cap = 3 * n;
  }
  if (ws.n_cap < n) {
    if (ws.qoff) cudaFree(ws.qoff);
    ws.qoff = nullptr;
    if (cudaMalloc(&ws.qoff, 3 * n * sizeof(uint64_t)) != cudaSuccess) return -1;
    ws.n_cap = n;
  }
  ws.soff = ws.qoff + ws.n_cap; ws.toff = ws.qoff + 2 * ws.n_cap;
  uint64_t qs = 0, ss = 0, tasks1 = 0, tasks2 = 0, tiles = 0;
  uint32_t max_q = 0;
  for (size_t i = 0; i < n; i++) {
    ws.h_off[i] = qs; ws.h_off[n + i] = ss; ws.h_off[2 * n + i] = tiles;
    tiles += wave_tb_tiles_of(h_qlens[i], h_tlens[i]);
    max_q = std::max(max_q, h_qlens[i]);
    const uint64_t tr = ((uint64_t)h_tlens[i] + 7u) >> 3;
    qs += ((uint64_t)h_qlens[i] + 7u) & ~7ull;
    ss += tr;
    tasks1 += (tr + 31) >> 5; tasks2 += (tr + 63) >> 6;
  }
  // Row-blocks of one pair form a pipeline that runs at the pace of its slowest stage, so the warp-workers must be
  // spread evenly over the SM sub-partitions: 16 rows per lane as long as that still gives every sub-partition a task
  // (fewer, longer pipeline stages), else 8; and never more workers than tasks (the hardware deals CTAs round-robin).
  const bool many = tasks2 * 4 >= (uint64_t)n_sm * 4 * 3;
  // the s16x2 kernel's row-block is 64 strips; with few tasks the finer int32 blocks (SR = 1) fill the machine better
  // GLOBAL: the s16x2 kernel wins once there are two row-blocks per scheduler (16k x 16k: 48 pairs and up, 3670 vs 2757 GCUPS at
  // 128); below that the finer int32 blocks keep more schedulers busy (32 pairs: 2521 vs 1964)
  const bool many16 = (algo == ALGO_GLOBAL) ? (tasks2 >= (uint64_t)n_sm * 8) : many;
  const bool tbm = (start == START_TB);
  if (tbm && !(s16_inout && *s16_inout && algo == ALGO_LOCAL)) return -1;     // the traceback variant exists for the s16x2 LOCAL kernel only
  const bool use_s16 = s16_inout && *s16_inout && (many16 || force_s16 || tbm);
  if (s16_inout) *s16_inout = use_s16;
  const int SR = (many || use_s16) ? 2 : 1;
  const uint64_t tasks = (SR == 2) ? tasks2 : tasks1;
  const uint32_t max_tr = (max_t + 7u) >> 3;
  const uint32_t maxblk = std::max<uint32_t>(1u, (max_tr + 32u * SR - 1) / (32u * SR));
  const size_t n_tickets = n * maxblk;
  if (n_tickets >= 0xFFFFFFF0ull) return -1;
  const int npass = (start == START_POS) ? 2 : 1;
  if (ws.bnd_cap < 2 * qs || (uint64_t)ws.next_tag + (uint64_t)npass * maxblk >= 0xFFFFFFF0ull) {
    if (ws.bnd_cap < 2 * qs) {
      if (ws.bnd) cudaFree(ws.bnd);
      ws.bnd = nullptr;
      if (cudaMalloc(&ws.bnd, std::max<uint64_t>(2 * qs, 16) * sizeof(uint4)) != cudaSuccess) return -1;
      ws.bnd_cap = std::max<uint64_t>(2 * qs, 16);
    }
    cudaMemsetAsync(ws.bnd, 0, ws.bnd_cap * sizeof(uint4), str);   // tag 0 = never written
    ws.next_tag = 1;
  }
  if (ws.cand_cap < ss) {
    if (ws.cand) cudaFree(ws.cand);
    ws.cand = nullptr;
    if (cudaMalloc(&ws.cand, std::max<uint64_t>(ss, 16) * sizeof(int2)) != cudaSuccess) return -1;
    ws.cand_cap = ss;
  }
  const size_t prog_words = n + 1;
  if (ws.prog_cap < prog_words) {
    if (ws.prog) cudaFree(ws.prog);
    ws.prog = nullptr;
    if (cudaMalloc(&ws.prog, prog_words * sizeof(uint32_t)) != cudaSuccess) return -1;
    ws.prog_cap = prog_words;
  }
  cudaMemcpyAsync(ws.qoff, ws.h_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, str);
  cudaMemcpyAsync(ws.soff, ws.h_off + n, n * sizeof(uint64_t), cudaMemcpyHostToDevice, str);
  if (tbm) cudaMemcpyAsync(ws.toff, ws.h_off + 2 * n, n * sizeof(uint64_t), cudaMemcpyHostToDevice, str);
  WaveArgs wa; wa.a = a; if (tbm) wa.a.tb_off = ws.toff; wa.algo = algo; wa.maxblk = maxblk; wa.qoff = ws.qoff; wa.soff = ws.soff;
  wa.bnd = ws.bnd; wa.bnd_stride = qs; wa.cand = ws.cand; wa.done = ws.prog; wa.ticket = ws.prog + n;
  wa.kshift = 0; wa.floor16 = 0; wa.mm = 0x00204081u; wa.sixteen = 16;
  if (use_s16 && algo == ALGO_GLOBAL && !wave16_global_range(a, max_q, max_t, &wa.kshift, &wa.floor16)) return -1;   // cannot happen: checked by wave16_eligible
  if (use_s16 && algo == ALGO_SEMI_GLOBAL && !wave16_semi_range(a, max_q, max_t, &wa.kshift, &wa.floor16)) return -1;
  const int ctas_per_sm = tbm ? 2 : (use_s16 ? WAVE16_MIN_CTAS : 4);
  // <= 16 (24) warp-workers per SM.  Once there is at least one CTA (= one warp per scheduler) for every SM, only whole multiples of
  // the SM count are started: a pipeline runs at the pace of its most crowded scheduler, so 1.7 workers per scheduler (two on
  // some, one on others) is slower than one everywhere and a second round of tickets (32 pairs of 16k x 16k: 1651 -> see DESIGN.md)
  uint64_t ctas = std::min<uint64_t>((uint64_t)n_sm * ctas_per_sm, (tasks + kWaveWarps - 1) / kWaveWarps);
  if (ctas > (uint64_t)n_sm) ctas = ctas / (uint64_t)n_sm * (uint64_t)n_sm;
  const int blocks = (int)std::max<uint64_t>(1, ctas);
  int launches = 0;
  for (int pass = 0; pass < npass; pass++) {
    cudaMemsetAsync(ws.prog, 0, prog_words * sizeof(uint32_t), str);
    wa.pass = pass; wa.tag_base = ws.next_tag; ws.next_tag += maxblk;
    if (use_s16 && algo == ALGO_GLOBAL) k_wave16_global<<<blocks, kWaveWarps * 32, 0, str>>>(wa);
    else if (use_s16 && algo == ALGO_SEMI_GLOBAL) k_wave16_semi<<<blocks, kWaveWarps * 32, 0, str>>>(wa);
    else if (use_s16 && tbm) k_wave16_local_t<true><<<blocks, kWaveWarps * 32, 0, str>>>(wa);
    else if (use_s16) k_wave16_local_t<false><<<blocks, kWaveWarps * 32, 0, str>>>(wa);
    else if (SR == 2) wave_launch_kernel<2>(algo, blocks, wa, str); else wave_launch_kernel<1>(algo, blocks, wa, str);
    launches++;
  }
  return launches;
}

}  // namespace gb200
