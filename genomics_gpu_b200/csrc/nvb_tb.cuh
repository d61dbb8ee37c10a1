// nvb_tb.cuh -- nvbio-style BANDED scoring and TRACEBACK batches (the rest of SURVEY.md 8f row f4); included by nvb_batched.cu.
//
// Reference (Non-CDP/NvB/nvbio/alignment):
//   BatchedBandedAlignmentScore<BAND_LEN, stream>::enact       batched.h:337-357, batched_banded_inl.h -> banded_alignment_score<BAND_LEN>
//                                                              gotoh/gotoh_banded_inl.h:405-627, sw/sw_banded_inl.h:341-512, ed/ed_banded_inl.h
//   BatchedAlignmentTraceback<CHECKPOINTS, stream>::enact      batched.h:436-451 -> alignment_traceback (alignment_inl.h:361-467),
//                                                              direction vectors gotoh/gotoh_inl.h:513-555, walkers gotoh_inl.h:1806-1872, sw/sw_inl.h:1650-1698
//   BatchedBandedAlignmentTraceback<BAND_LEN, CHECKPOINTS, ..> batched.h:463-478 -> banded_alignment_traceback (banded_inl.h:363-427),
//                                                              walkers gotoh_banded_inl.h:860-927, sw_banded_inl.h:735-790
// The reference keeps short2 checkpoints every CHECKPOINTS rows and recomputes one 4-bit "flow submatrix" window at a time in local
// memory, one thread per job walking back through the windows.  Here a B200 has the HBM to keep the whole flow matrix of a chunk of
// jobs (0.5 B per cell, [word][job] so that a warp's stores are coalesced): ONE forward sweep per job writes every direction nibble,
// and a second tiny kernel walks them -- no recomputation, no checkpoints.  Same arithmetic, same direction rules, same walker state
// machine: bit-exact inside the reference's int16 checkpoint domain (oracle/nvbio_compat.c is the pinned restatement).
//
// One Gotoh-shaped code path serves the three aligner families: Smith-Waterman / edit distance are the recurrences with open == extend,
// for which F == H(top) + G and E == H(left) + I, the E/F "extended" bits are never set and the Gotoh walker degenerates to the
// Smith-Waterman one (argued in DESIGN.md K7, checked against the reference's separate SW code by the oracle tests).
#pragma once

namespace nvb200 {

enum { DV_SUB = 0, DV_INS = 1, DV_DEL = 2, DV_SINK = 3, DV_IEXT = 4, DV_DEXT = 8 };     // aln::DirectionVector (alignment_base.h:139-150)

struct TbArgs {
  uint32_t* dirs; size_t slots;            // flow words: dirs[word * slots + slot], slot = job - first of the chunk
  uint32_t rows_cap;                       // full matrix: word = band8 * rows_cap + (text row - 1);  banded: word = pattern row * W + (entry >> 3)
  int f_o, f_e, e_o, e_e, init_o, init_e;  // banded: costs of the pattern-consuming move (F, "top"), the text-consuming one (E, "left"), row-zero init
  int no_sink;                             // banded Smith-Waterman / edit distance: LOCAL cells are never marked SINK (sw_banded_inl.h:270-278)
  int32_t* src_i; int32_t* src_j; uint32_t* n_ops; uint8_t* ops; uint32_t ops_stride;
};

// ---- full-matrix forward sweep with direction nibbles ------------------------------------------------------------------------
// k_nvb_i32's band-of-8 sweep (same boundaries, same BestSink order) + per cell: hdir | edir | fdir (gotoh_inl.h:513-555), eight
// nibbles = one word per (band, text row).
template <int TYPE>
__global__ void __launch_bounds__(kThreads) k_nvb_tb_fwd(const NvbArgs a, const TbArgs tb) {
  const uint32_t idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.count) return;
  const uint32_t job = a.first + idx;
  const Job jb = load_job(a, job);
  const uint8_t* __restrict__ P = a.pbytes + a.poffs[jb.pi];
  const uint8_t* __restrict__ T = a.tbytes + a.toffs[jb.ti];
  const int M = (int)jb.M, N = (int)jb.N;
  uint2* __restrict__ col = a.scratch + (size_t)blockIdx.x * a.rows_cap * kThreads + threadIdx.x;
  const uint32_t stride = (uint32_t)kThreads;
  const int go_t = a.go_t, ge_t = a.ge_t, go_p = a.go_p, ge_p = a.ge_p, match = a.match, mismatch = a.mismatch;
  uint32_t* __restrict__ dw = tb.dirs + idx;

  int best = INT_MIN, bi = -1, bj = -1;
  for (int block = 0; block < M; block += 8) {
    int q[8];
#pragma unroll
    for (int j = 0; j < 8; j++) q[j] = (block + j < M) ? (int)P[block + j] : 256 + j;
    int H[9], F[9];
#pragma unroll
    for (int j = 0; j <= 8; j++) {
      const int c = block + j;
      H[j] = (TYPE == NVBX_LOCAL || c == 0) ? 0 : go_p + ge_p * (c - 1);
      F[j] = kNeg32;
    }
    const bool first = block == 0, last = block + 8 >= M;
    const int mj = last ? M - block : 8;
    uint2 nxt = first ? make_uint2(0, 0) : col[0];
    uint32_t* wp = dw + (size_t)(block >> 3) * tb.rows_cap * tb.slots;
    for (int i = 1; i <= N; i++) {
      int hl, e;
      if (first) { hl = (TYPE == NVBX_GLOBAL) ? go_t + ge_t * (i - 1) : 0; e = (TYPE == NVBX_LOCAL) ? 0 : kNeg32; }
      else { hl = (int)nxt.x; e = (int)nxt.y; if (i < N) nxt = col[(size_t)i * stride]; }
      const int t = (int)T[i - 1];
      int diag = H[0];
      H[0] = hl;
      uint32_t w = 0;
#pragma unroll
      for (int j = 1; j <= 8; j++) {
        const int ftop = F[j] + ge_t, htop = H[j] + go_t;
        F[j] = max(ftop, htop);
        const int eleft = e + ge_p, hleft = H[j - 1] + go_p;
        e = max(eleft, hleft);
        const int d = diag + ((t == q[j - 1]) ? match : mismatch);
        const int top = F[j], left = e;
        int h = max(max(left, top), d);
        if (TYPE == NVBX_LOCAL) h = max(h, 0);
        uint32_t nib = top > left ? (top > d ? DV_DEL : DV_SUB) : (left > d ? DV_INS : DV_SUB);
        if (TYPE == NVBX_LOCAL && h == 0) nib = DV_SINK;
        nib |= (eleft > hleft ? DV_IEXT : 0u) | (ftop > htop ? DV_DEXT : 0u);
        w |= nib << (4 * (j - 1));
        diag = H[j];
        H[j] = h;
      }
      wp[(size_t)(i - 1) * tb.slots] = w;
      if (!last) col[(size_t)(i - 1) * stride] = make_uint2((uint32_t)H[8], (uint32_t)e);
      if (TYPE == NVBX_LOCAL) {
        int rowmax = INT_MIN;
#pragma unroll
        for (int j = 1; j <= 8; j++) if (j <= mj) rowmax = max(rowmax, H[j]);
        if (rowmax >= best) {
          int cj = 0;
#pragma unroll
          for (int j = 1; j <= 8; j++) if (j <= mj && H[j] == rowmax) cj = j;
          if (rowmax > best || reported_later(i, block + cj, bi, bj, a.band_shift)) { best = rowmax; bi = i; bj = block + cj; }
        }
      } else if (last) {
        const int hM = select_reg(H, mj);
        if (TYPE == NVBX_SEMI_GLOBAL) { if (hM >= best) { best = hM; bi = i; bj = M; } }
        else if (i == N) { best = hM; bi = N; bj = M; }
      }
    }
  }
  a.score[job] = best; a.sink_i[job] = bi; a.sink_j[job] = bj;
}

// the walker of alignment_inl.h:361-467 + gotoh_inl.h:1806-1872 over the stored flow words; ops in push order (end -> source)
__global__ void __launch_bounds__(kThreads) k_nvb_tb_walk(const NvbArgs a, const TbArgs tb, int type) {
  const uint32_t idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.count) return;
  const uint32_t job = a.first + idx;
  int row = a.sink_i[job], col = a.sink_j[job];
  uint8_t* __restrict__ out = tb.ops + (size_t)job * tb.ops_stride;
  uint32_t n = 0;
  if (row < 0 || col < 0) { tb.src_i[job] = -1; tb.src_j[job] = -1; tb.n_ops[job] = 0; return; }
  const uint32_t* __restrict__ dw = tb.dirs + idx;
  int state = 0;                                                   // HSTATE / ESTATE / FSTATE
  int wrow = -1, wband = -1; uint32_t w = 0;
  while (row > 0 && col > 0) {
    const int b = (col - 1) >> 3;
    if (b != wband || row != wrow) { w = dw[((size_t)b * tb.rows_cap + (size_t)(row - 1)) * tb.slots]; wband = b; wrow = row; }
    const uint32_t op = (w >> (4 * ((col - 1) & 7))) & 15u, h_op = op & 3u;
    if (type == NVBX_LOCAL && state == 0 && h_op == DV_SINK) break;
    if (state == 1) { if ((op & DV_IEXT) == 0) state = 0; --col; out[n++] = DV_INS; }
    else if (state == 2) { if ((op & DV_DEXT) == 0) state = 0; --row; out[n++] = DV_DEL; }
    else if (h_op == DV_INS) state = 1;
    else if (h_op == DV_DEL) state = 2;
    else { --col; --row; out[n++] = DV_SUB; }
  }
  if (type != NVBX_LOCAL && row == 0) for (; col > 0; --col) out[n++] = DV_INS;      // first row / column are not stored (alignment_inl.h:444-459)
  if (type == NVBX_GLOBAL && col == 0) for (; row > 0; --row) out[n++] = DV_DEL;
  tb.src_i[job] = row; tb.src_j[job] = col; tb.n_ops[job] = n;
}

// ---- banded DP: row i = pattern symbol i, entry j = text symbol i + j (gotoh_banded_inl.h:405-627) ----------------------------
// One thread per job, the band (H, F) and the text window in registers -- no scratch at all.  PACKED: the reference's generic text
// cache keeps 2 bits per symbol (alignment_base_inl.h:75-79; BAND_LEN 31), so cached symbols are truncated to their low two bits,
// the 255 sentinel past the end of the text included.  DIRS: also write the flow nibbles (banded traceback).
template <int B, int TYPE, bool DIRS>
__global__ void __launch_bounds__(kThreads) k_nvb_banded(const NvbArgs a, const TbArgs tb, int want_sinks) {
  constexpr bool PACKED = !(B == 3 || B == 5 || B == 7 || B == 15);
  constexpr int W = (B + 7) / 8;
  const uint32_t idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.count) return;
  const uint32_t job = a.first + idx;
  const Job jb = load_job(a, job);
  const uint8_t* __restrict__ P = a.pbytes + a.poffs[jb.pi];
  const uint8_t* __restrict__ T = a.tbytes + a.toffs[jb.ti];
  const int M = (int)jb.M, N = (int)jb.N;
  if (N < M) {                                                      // the reference returns false and reports nothing: the BestSink stays as constructed
    a.score[job] = -(1 << 30); a.sink_i[job] = -1; a.sink_j[job] = -1;      // BestSink(): Field_traits<int32>::min(), numbers.h:832-836
    return;
  }
  const int f_o = tb.f_o, f_e = tb.f_e, e_o = tb.e_o, e_e = tb.e_e, match = a.match, mismatch = a.mismatch;
  int H[B], F[B], c[B - 1];
#pragma unroll
  for (int j = 0; j < B - 1; j++) { const int g = (j < N) ? (int)T[j] : 255; c[j] = PACKED ? (g & 3) : g; }   // (the reference reads these unguarded)
  H[0] = 0;
#pragma unroll
  for (int j = 1; j < B; j++) H[j] = (TYPE == NVBX_GLOBAL) ? tb.init_o + (j - 1) * tb.init_e : 0;
#pragma unroll
  for (int j = 0; j < B; j++) F[j] = kNeg32;
  int best = INT_MIN, bx = -1, by = -1;
  uint32_t* __restrict__ dw = DIRS ? tb.dirs + idx : nullptr;
  const bool mark_sink = (TYPE == NVBX_LOCAL) && !tb.no_sink;
  for (int i = 0; i < M; i++) {
    const int q = (int)P[i];
    const int gl = (i + B - 1 < N) ? (int)T[i + B - 1] : 255;
    uint32_t w[W];
#pragma unroll
    for (int k = 0; k < W; k++) w[k] = 0;
    int E; uint32_t edir = 0;
    {   // entry 0: no left neighbour
      const int ftop = F[1] + f_e, htop = H[1] + f_o;
      F[0] = max(ftop, htop);
      const int d = H[0] + ((c[0] == q) ? match : mismatch), top = F[0];
      int h = max(top, d);
      if (TYPE == NVBX_LOCAL) h = max(h, 0);
      H[0] = h;
      if (DIRS) {
        uint32_t nib = (top > d) ? DV_INS : DV_SUB;
        if (mark_sink && h == 0) nib = DV_SINK;
        w[0] |= nib | (ftop > htop ? DV_DEXT : 0u);
      }
      E = h + e_o;
    }
#pragma unroll
    for (int j = 1; j < B - 1; j++) {
      const int ftop = F[j + 1] + f_e, htop = H[j + 1] + f_o;
      F[j] = max(ftop, htop);
      const int g = c[j]; c[j - 1] = g;
      const int d = H[j] + ((g == q) ? match : mismatch), top = F[j], left = E;
      int h = max(max(top, left), d);
      if (TYPE == NVBX_LOCAL) h = max(h, 0);
      H[j] = h;
      if (DIRS) {
        uint32_t nib = top > left ? (top > d ? DV_INS : DV_SUB) : (left > d ? DV_DEL : DV_SUB);
        if (mark_sink && h == 0) nib = DV_SINK;
        nib |= edir | (ftop > htop ? DV_DEXT : 0u);
        w[j >> 3] |= nib << (4 * (j & 7));
      }
      const int eleft = E + e_e, ediag = h + e_o;
      edir = (eleft > ediag) ? DV_IEXT : 0u;
      E = max(ediag, eleft);
    }
    c[B - 2] = PACKED ? (gl & 3) : gl;
    {   // entry B - 1: nothing above (F = -inf); compares the UNPACKED symbol (gotoh_banded_inl.h:568-577)
      F[B - 1] = kNeg32;
      const int d = H[B - 1] + ((gl == q) ? match : mismatch), left = E;
      int h = max(left, d);
      if (TYPE == NVBX_LOCAL) h = max(h, 0);
      H[B - 1] = h;
      if (DIRS) {
        uint32_t nib = (left > d) ? DV_DEL : DV_SUB;
        if (mark_sink && h == 0) nib = DV_SINK;
        nib |= edir;
        w[(B - 1) >> 3] |= nib << (4 * ((B - 1) & 7));
      }
    }
    if (DIRS) {
#pragma unroll
      for (int k = 0; k < W; k++) dw[((size_t)i * W + k) * tb.slots] = w[k];
    }
    if (TYPE == NVBX_LOCAL) {                                       // every cell is reported, row by row, entry by entry; the LAST maximum wins
      int rm = H[0];
#pragma unroll
      for (int j = 1; j < B; j++) rm = max(rm, H[j]);
      if (rm >= best) {
        int cj = 0;
#pragma unroll
        for (int j = 1; j < B; j++) if (H[j] == rm) cj = j;
        best = rm; bx = i + cj + 1; by = i + 1;
      }
    }
  }
  if (TYPE == NVBX_GLOBAL) { best = H[B - 1]; bx = M + B - 1; by = M; }
  else if (TYPE == NVBX_SEMI_GLOBAL) {
    const int lim = min(M + B - 1, N) - (M - 1);
    best = H[0]; bx = M; by = M;
#pragma unroll
    for (int j = 1; j < B; j++) if (j < lim && H[j] >= best) { best = H[j]; bx = M + j; by = M; }
  }
  a.score[job] = best;
  if (want_sinks) { a.sink_i[job] = bx; a.sink_j[job] = by; }
}

// the walkers of gotoh_banded_inl.h:860-927 / sw_banded_inl.h:735-790 (the latter is the former with open == extend, see the header)
template <int B>
__global__ void __launch_bounds__(kThreads) k_nvb_banded_walk(const NvbArgs a, const TbArgs tb, int type) {
  constexpr int W = (B + 7) / 8;
  const uint32_t idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= a.count) return;
  const uint32_t job = a.first + idx;
  const int sx = a.sink_i[job], sy = a.sink_j[job];
  uint8_t* __restrict__ out = tb.ops + (size_t)job * tb.ops_stride;
  if (sx < 0 || sy < 0) { tb.src_i[job] = -1; tb.src_j[job] = -1; tb.n_ops[job] = 0; return; }
  const uint32_t* __restrict__ dw = tb.dirs + idx;
  int entry = sx - sy, row = sy - 1, state = 0;
  uint32_t n = 0;
  bool found = false;
  while (row >= 0 && entry >= 0 && entry < B) {                     // (the reference asserts 0 <= entry < BAND_LEN)
    const uint32_t w = dw[((size_t)row * W + (entry >> 3)) * tb.slots];
    const uint32_t op = (w >> (4 * (entry & 7))) & 15u, h_op = op & 3u;
    if (type == NVBX_LOCAL && state == 0 && h_op == DV_SINK) { found = true; break; }
    if (state == 1) { if ((op & DV_IEXT) == 0) state = 0; --entry; out[n++] = DV_DEL; }
    else if (state == 2) { if ((op & DV_DEXT) == 0) state = 0; ++entry; --row; out[n++] = DV_INS; }
    else if (h_op == DV_DEL) state = 1;
    else if (h_op == DV_INS) state = 2;
    else { --row; out[n++] = DV_SUB; }
  }
  const int y = found ? row + 1 : 0;
  tb.src_i[job] = entry + y; tb.src_j[job] = y; tb.n_ops[job] = n;
}

}  // namespace nvb200
