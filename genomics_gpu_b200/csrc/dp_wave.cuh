// dp_wave.cuh -- K3: the intra-task (warp-cooperative) anti-diagonal wavefront kernel for LONG pairs.
//
// One pair is cut into ROW-BLOCKS of 32 lanes x 8*SR target rows.  A row-block is one warp-task: lane L owns SR
// consecutive 8-row strips and walks the query left to right, one column per step, skewed by one column per lane
// (lane L is on column k - L at step k), so lane L's bottom row (H, E) reaches lane L+1 with __shfl_up_sync one step
// later -- no shared-memory column array at all inside a block.  Lane 31's bottom row is the top boundary of the next
// row-block of the same pair; it goes through a small per-warp shared-memory ring and is flushed 32 columns at a time
// (coalesced) to a per-pair double buffer in global memory (L2-resident), with a release/acquire progress counter
// per task.  Tasks are handed out by a global ticket counter in (pair, block) order, so a task only ever waits on a
// LOWER ticket that some resident warp already holds: deadlock-free for any grid size, no cooperative launch, and all
// row-blocks of a pair (and of the next pairs) run concurrently as a pipeline across the whole GPU.
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
  uint2* bnd;                // two column buffers: bnd[buf * bnd_stride + qoff[p] + c] = (H, E) below the block's last row
  uint64_t bnd_stride;
  int2* cand;                // per strip: (score, position) candidate of the strip
  uint32_t* prog;            // per ticket: number of columns whose bottom boundary is published
  uint32_t* ticket;          // global ticket counter
  uint32_t* done;            // per pair: finished row-blocks
};

struct WaveGeom {
  int ncol;        // query columns swept
  int nstrips;     // target strips swept (pass-local numbering)
  int strip0;      // first strip in the sequence numbering (semi-global start pass begins at gend_reg)
  int qwords, twords;   // local start pass: rend_reg, gend_reg
  int fwd;         // forward score (start passes)
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
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

template <int ALGO, int SR>
__global__ void __launch_bounds__(kWaveWarps * 32) k_wave(const WaveArgs w) {
  constexpr int ROWS = 8 * SR;
  constexpr bool SEMI = (ALGO == ALGO_SEMI_GLOBAL), LOCALA = (ALGO == ALGO_LOCAL), GLOB = (ALGO == ALGO_GLOBAL);
  __shared__ uint4 s_in[kWaveWarps][32];     // top boundary + selector of the current 32-column chunk: (H, E, sel, -)
  __shared__ uint2 s_out[kWaveWarps][32];    // bottom boundary of the chunk being produced
  const DpArgs& a = w.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint4* sin = s_in[wib]; uint2* sout = s_out[wib];
  const Sc sc = make_sc(a, GLOB);
  const int oe = sc.oe, ext = sc.ext;
  const bool top_zero = SEMI && (a.head == SKIP_BOTH);
  const uint32_t total = (uint32_t)a.n * w.maxblk;

  for (;;) {
    uint32_t T = 0;
    if (lane == 0) T = atomicAdd(w.ticket, 1u);
    T = __shfl_sync(0xFFFFFFFFu, T, 0);
    if (T >= total) break;
    const uint32_t pair = T / w.maxblk, blk = T - pair * w.maxblk;
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
    const bool last_blk = ((int)blk == nblk - 1);

    // ---- per-row substitution tables (4 bytes: score against query A, C, G, T) and left-boundary state ----
    uint32_t tword[SR];
    uint32_t tbl[ROWS];
    int p[ROWS], f[ROWS], gl[ROWS];
#pragma unroll
    for (int s = 0; s < SR; s++) {
      tword[s] = (lstrip + s < g.nstrips) ? wave_tword(w, g, tw, tlen, lstrip + s) : 0u;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const uint32_t tc = nib(tword[s], r);
        const uint32_t b0 = (uint32_t)score_general(sc, 1u, tc) & 0xFFu, b1 = (uint32_t)score_general(sc, 3u, tc) & 0xFFu;
        const uint32_t b2 = (uint32_t)score_general(sc, 7u, tc) & 0xFFu, b3 = (uint32_t)score_general(sc, 4u, tc) & 0xFFu;
        tbl[8 * s + r] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        const int t = 8 * (g.strip0 + lstrip + s) + r;          // row index in the pass' sequence numbering
        if (LOCALA) { p[8 * s + r] = 0; f[8 * s + r] = 0; }
        else if (GLOB) { p[8 * s + r] = (t == 0) ? 0 : -(sc.o + ext * t); f[8 * s + r] = kMinusInf; }     // global.h:63-71
        else { p[8 * s + r] = 0; f[8 * s + r] = kMinusInf; }                                              // free left column (:114-121)
        gl[8 * s + r] = -oe;
      }
    }
    const uint32_t ntab = (uint32_t)sc.nscore & 0xFFu;          // query N under the LOCAL rule: same score in every row

    // ---- tracking state ----
    int best[SR], bpos[SR];
#pragma unroll
    for (int s = 0; s < SR; s++) { best[s] = LOCALA ? 0 : kMinusInf; bpos[s] = SEMI ? ((w.pass == 0) ? qlen : 0) : 0; }

    const uint2* bnd_in = w.bnd + (uint64_t)((blk + 1) & 1) * w.bnd_stride + w.qoff[pair];
    uint2* bnd_out = w.bnd + (uint64_t)(blk & 1) * w.bnd_stride + w.qoff[pair];
    const uint32_t* prog_in = w.prog + (T - 1);
    uint32_t* prog_out = w.prog + T;

    int h_bot = 0, e_bot = 0; uint32_t sel_cur = 0;
    const int nsteps = ncol + 31;
    for (int k0 = 0; k0 < nsteps; k0 += 32) {
      // ---- chunk prologue: top boundary and selectors of columns [k0, k0 + 32) ----
      if (k0 < ncol) {
        const int c = k0 + lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (blk == 0) {
          if (c < ncol) {
            if (LOCALA) { v.x = 0; v.y = 0; }
            else if (GLOB) { v.x = (uint32_t)(c == 0 ? 0 : -(sc.o + ext * c)); v.y = (uint32_t)kMinusInf; }      // global.h:57-60
            else {
              // semi-global hands E over as the first row's own E(0,c) = max(H(-1,c) - oe, E(-1,c) - e)
              const int htop = top_zero ? 0 : (c == 0 ? 0 : -(sc.o + ext * c));                                  // :87-99
              const int etop = top_zero ? 0 : kMinusInf;
              v.x = (uint32_t)htop; v.y = (uint32_t)max(htop - oe, etop - ext);
            }
          }
        } else {
          const uint32_t need = (uint32_t)min(k0 + 32, ncol);
          if (lane == 0) { while (ld_acquire_u32(prog_in) < need) __nanosleep(40); }
          __syncwarp();
          if (c < ncol) { const uint2 t2 = __ldcg(bnd_in + c); v.x = t2.x; v.y = t2.y; }
        }
        if (c < ncol) {
          const uint32_t qword = wave_qword(w, g, qw, qlen, c >> 3);
          const uint32_t qc = nib(qword, c & 7);
          const uint32_t idx = base_idx(qc);
          if (idx < 4) v.z = idx * 0x1111u + 0x8880u;
          else if (!GLOB && qc == sc.nval) v.z = 0xCCC4u;       // byte 4 = ntab
          else v.z = kWaveExotic | (qc << 20);
        }
        __syncwarp();
        sin[lane] = v;
        __syncwarp();
      }
      const int kend = min(k0 + 32, nsteps);
      for (int k = k0; k < kend; k++) {
        // ---- receive (H, E) of the row above and the column's selector from the lane above ----
        int h_in = __shfl_up_sync(0xFFFFFFFFu, h_bot, 1);
        int e_in = __shfl_up_sync(0xFFFFFFFFu, e_bot, 1);
        uint32_t sel = __shfl_up_sync(0xFFFFFFFFu, sel_cur, 1);
        if (lane == 0) { const uint4 t4 = sin[k & 31]; h_in = (int)t4.x; e_in = (int)t4.y; sel = t4.z; }
        const int c = k - lane;
        if (c >= 0 && c < ncol) {
          sel_cur = sel;
          int sv[ROWS];
          if (sel & kWaveExotic) {
            const uint32_t qc = (sel >> 20) & 15u;
#pragma unroll
            for (int s = 0; s < SR; s++)
#pragma unroll
              for (int r = 0; r < 8; r++) sv[8 * s + r] = score_general(sc, qc, nib(tword[s], r));
          } else {
#pragma unroll
            for (int r = 0; r < ROWS; r++) sv[r] = (int)prmt(tbl[r], ntab, sel);
          }
          int up = h_in, e = e_in;
          int hh[ROWS];
#pragma unroll
          for (int r = 0; r < ROWS; r++) {
            const int d = p[r] + sv[r];
            int h;
            if (SEMI) {
              f[r] = __viaddmax_s32(f[r], -ext, gl[r]);               // F(t,q) = max(H(t,q-1) - oe, F(t,q-1) - e)
              const int m = max(d, f[r]);
              h = max(m, e);
              e = __viaddmax_s32(e, -ext, m - oe);                    // E(t+1,q) = max(max(D,F) - oe, E - e)  (oe >= e)
              gl[r] = h - oe;
            } else {
              h = LOCALA ? __vimax3_s32_relu(d, f[r], e) : __vimax3_s32(d, f[r], e);
              const int dm = d - oe;
              f[r] = __viaddmax_s32(f[r], -ext, dm);
              e = __viaddmax_s32(e, -ext, dm);
            }
            hh[r] = h; p[r] = up; up = h;
          }
          h_bot = up; e_bot = e;
          // ---- tracking ----
          if (LOCALA) {
#pragma unroll
            for (int s = 0; s < SR; s++) {
              const int cm = max(__vimax3_s32(hh[8 * s], hh[8 * s + 1], hh[8 * s + 2]),
                                 __vimax3_s32(__vimax3_s32(hh[8 * s + 3], hh[8 * s + 4], hh[8 * s + 5]), hh[8 * s + 6], hh[8 * s + 7]));
              if (cm > best[s]) {
                int rr = 7;
#pragma unroll
                for (int r = 6; r >= 0; r--) rr = (hh[8 * s + r] == cm) ? r : rr;
                best[s] = cm; bpos[s] = (c << 3) | rr;
              }
            }
          } else if (c == qlen - 1) {
            if (SEMI) {
#pragma unroll
              for (int s = 0; s < SR; s++)
#pragma unroll
                for (int r = 0; r < 8; r++) {
                  const int t = 8 * (g.strip0 + lstrip + s) + r;
                  if (t < tlen && hh[8 * s + r] > best[s]) { best[s] = hh[8 * s + r]; bpos[s] = t; }
                }
            } else {
              const int t_last = tlen - 1;
#pragma unroll
              for (int s = 0; s < SR; s++)
                if ((t_last >> 3) == lstrip + s) {
                  int v = hh[8 * s];
#pragma unroll
                  for (int r = 1; r < 8; r++) v = ((t_last & 7) == r) ? hh[8 * s + r] : v;
                  a.score[pair] = v;                                  // H(ref_len-1, read_len-1)  (global.h:285-299)
                }
            }
          }
          if (lane == 31) sout[c & 31] = make_uint2((uint32_t)h_bot, (uint32_t)e_bot);
        }
        // ---- publish the bottom boundary 32 columns at a time ----
        const int c31 = k - 31;
        if (!last_blk && c31 >= 0 && (((c31 & 31) == 31) || c31 == ncol - 1)) {
          __syncwarp();
          const int cb = (c31 & ~31) + lane;
          if (cb <= c31) __stcg(bnd_out + cb, sout[lane]);
          __syncwarp();
          if (lane == 0) { __threadfence(); st_release_u32(prog_out, (uint32_t)(c31 + 1)); }
        }
      }
    }
    // ---- per-strip candidates, then the last block of the pair folds them ----
    if (!GLOB) {
      int2* cand = w.cand + w.soff[pair];
#pragma unroll
      for (int s = 0; s < SR; s++)
        if (lstrip + s < g.nstrips) __stcg(&cand[lstrip + s], make_int2(best[s], bpos[s]));
      __syncwarp();
      uint32_t old = 0;
      if (lane == 0) { __threadfence(); old = atomicAdd(w.done + pair, 1u); }
      old = __shfl_sync(0xFFFFFFFFu, old, 0);
      if ((int)old == nblk - 1) { __threadfence(); wave_finalize<ALGO>(w, pair, g, lane); }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct WaveScratch {
  uint64_t* qoff = nullptr; uint64_t* soff = nullptr; size_t n_cap = 0;
  uint64_t* h_off = nullptr; size_t h_cap = 0;                    // pinned staging: qoff[n] then soff[n]
  uint2* bnd = nullptr; size_t bnd_cap = 0;
  int2* cand = nullptr; size_t cand_cap = 0;
  uint32_t* prog = nullptr; size_t prog_cap = 0;                  // prog[...], then done[n], then ticket
  void release() {
    if (qoff) cudaFree(qoff);
    if (h_off) cudaFreeHost(h_off);
    if (bnd) cudaFree(bnd);
    if (cand) cudaFree(cand);
    if (prog) cudaFree(prog);
    qoff = soff = nullptr; h_off = nullptr; bnd = nullptr; cand = nullptr; prog = nullptr; n_cap = h_cap = bnd_cap = cand_cap = prog_cap = 0;
  }
};

inline bool wave_eligible(const DpArgs& a, int algo, int start, bool second, uint32_t min_q, uint32_t min_t) {
  if (second || start == START_TB) return false;
  if (min_q == 0 || min_t == 0) return false;
  if (a.match < 0 || a.match > 127 || a.mismatch < 0 || a.mismatch > 127) return false;
  if (a.n_score < -127) return false;
  if (a.gap_open < 0 || a.gap_extend < 0) return false;
  if (algo == ALGO_LOCAL) return true;
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
inline int wave_launch(const DpArgs& a, int algo, int start, WaveScratch& ws, const uint32_t* h_qlens, const uint32_t* h_tlens,
                       uint32_t max_t, int n_sm, cudaStream_t str) {
  const size_t n = (size_t)a.n;
  if (ws.h_cap < 2 * n) {
    if (ws.h_off) cudaFreeHost(ws.h_off);
    ws.h_off = nullptr;
    if (cudaHostAlloc(&ws.h_off, 2 * n * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess) return -1;
    ws.h_cap = 2 * n;
  }
  if (ws.n_cap < n) {
    if (ws.qoff) cudaFree(ws.qoff);
    ws.qoff = nullptr;
    if (cudaMalloc(&ws.qoff, 2 * n * sizeof(uint64_t)) != cudaSuccess) return -1;
    ws.soff = ws.qoff + n; ws.n_cap = n;
  }
  ws.soff = ws.qoff + ws.n_cap;
  uint64_t qs = 0, ss = 0;
  for (size_t i = 0; i < n; i++) {
    ws.h_off[i] = qs; ws.h_off[n + i] = ss;
    qs += ((uint64_t)h_qlens[i] + 7u) & ~7ull;
    ss += ((uint64_t)h_tlens[i] + 7u) >> 3;
  }
  // rows per lane: 16 when that still leaves every SM several row-blocks to work on, else 8
  const uint64_t strips_total = ss;
  const int SR = (strips_total / 64 >= (uint64_t)n_sm * 8) ? 2 : 1;
  const uint32_t max_tr = (max_t + 7u) >> 3;
  const uint32_t maxblk = std::max<uint32_t>(1u, (max_tr + 32u * SR - 1) / (32u * SR));
  const size_t n_tickets = n * maxblk;
  if (n_tickets >= 0xFFFFFFF0ull) return -1;
  if (ws.bnd_cap < 2 * qs) {
    if (ws.bnd) cudaFree(ws.bnd);
    ws.bnd = nullptr;
    if (cudaMalloc(&ws.bnd, std::max<uint64_t>(2 * qs, 16) * sizeof(uint2)) != cudaSuccess) return -1;
    ws.bnd_cap = 2 * qs;
  }
  if (ws.cand_cap < ss) {
    if (ws.cand) cudaFree(ws.cand);
    ws.cand = nullptr;
    if (cudaMalloc(&ws.cand, std::max<uint64_t>(ss, 16) * sizeof(int2)) != cudaSuccess) return -1;
    ws.cand_cap = ss;
  }
  const size_t prog_words = n_tickets + n + 1;
  if (ws.prog_cap < prog_words) {
    if (ws.prog) cudaFree(ws.prog);
    ws.prog = nullptr;
    if (cudaMalloc(&ws.prog, prog_words * sizeof(uint32_t)) != cudaSuccess) return -1;
    ws.prog_cap = prog_words;
  }
  cudaMemcpyAsync(ws.qoff, ws.h_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, str);
  cudaMemcpyAsync(ws.soff, ws.h_off + n, n * sizeof(uint64_t), cudaMemcpyHostToDevice, str);
  WaveArgs wa; wa.a = a; wa.algo = algo; wa.maxblk = maxblk; wa.qoff = ws.qoff; wa.soff = ws.soff;
  wa.bnd = ws.bnd; wa.bnd_stride = qs; wa.cand = ws.cand; wa.prog = ws.prog; wa.done = ws.prog + n_tickets; wa.ticket = ws.prog + n_tickets + n;
  const int blocks = n_sm * 4;                                   // 16 resident warp-workers per SM
  int launches = 0;
  const int npass = (start == START_POS) ? 2 : 1;
  for (int pass = 0; pass < npass; pass++) {
    cudaMemsetAsync(ws.prog, 0, prog_words * sizeof(uint32_t), str);
    wa.pass = pass;
    if (SR == 2) wave_launch_kernel<2>(algo, blocks, wa, str); else wave_launch_kernel<1>(algo, blocks, wa, str);
    launches++;
  }
  return launches;
}

}  // namespace gb200
