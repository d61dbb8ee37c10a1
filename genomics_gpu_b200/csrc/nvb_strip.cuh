// nvb_strip.cuh -- nvbio-style batched SCORING (include/nvbio_b200.h, K7) on the K2 quad-strip layout of dp_s16.cuh.
//
// The first packed K7 kernel (k_nvb_s16, nvb_batched.cu) follows the reference's own blocking (gotoh/gotoh_inl.h:473-900: bands of
// pattern columns in registers, the band border (H, E) parked once per TEXT row) and therefore streams 8 B per 32 cells through a
// scratch in HBM: 1.4 GB written + read per 65536 jobs, two orders of magnitude more than the strings.  Here the blocking is K2's:
// one thread = one job, four consecutive 8-row strips of the TEXT in the four 16-bit lanes of two registers, and the only parked
// state is one (H, E) word + one selector byte per PATTERN column in shared memory (5 B x (M8 + 8) per job) -- nothing but the
// strings and 4 B of score per job ever touches HBM.
//
// Recurrences (alignment.h:456-530 -> gotoh_inl.h:513-555; restated in oracle/nvbio_compat.c): H-based gap opening for all three
// alignment types, a gap of length k costs Go + (k - 1) Ge.  With GASAL2's names oe = -Go, ext = -Ge, open = oe - ext:
//   GLOBAL / SEMI_GLOBAL : nc_sweep4<ALGO_SEMI_GLOBAL> (biased frame, E off the H chain), nvbio's boundaries
//                          H(0, j) = Go + Ge (j - 1), H(i, 0) = 0 (SEMI_GLOBAL) or Go + Ge (i - 1) (GLOBAL), E(i, 0) = F(0, j) = -inf;
//                          SEMI_GLOBAL scans column M (BestSink: the LAST best row wins, sink_inl.h:52-60), GLOBAL captures (N, M)
//   LOCAL                : nvb_local_sweep4 (clamped at 0, chain formulation, unbiased frame); the score is the running maximum of H
//                          (pad columns score 0 / -1 on the diagonal, pad rows the mismatch: they never exceed a true cell)
// LOCAL sinks keep the reference's band-major tie order (k_nvb_s16 / k_nvb_i32w); this kernel serves scores, and the sinks of
// GLOBAL / SEMI_GLOBAL.  Jobs with a symbol outside A/C/G/T go to the int32 worklist exactly as in k_nvb_s16.
#pragma once
#include <climits>
#include "dp_s16.cuh"

namespace gb200 {

__device__ __forceinline__ uint32_t nvb_row_table(uint32_t w, int sh, bool inside, uint32_t mis4, uint32_t delta) {
  return inside ? (mis4 ^ (delta << (8u * ((w >> sh) & 3u)))) : mis4;
}

// Score-only LOCAL sweep (H-based opening, clamped at 0): the chain formulation in an unbiased frame, without windows, keys or
// records.  Everything that travels -- the parked / handed-on column state, the left neighbours, the diagonal inputs -- is kept as
// G = H - oe, and the substitution tables hold s + oe, so that per cell pair there are exactly
//     ALU pipe : PRMT (s + oe), VIADDMNMX.RELU (F = max(F - e, G(left), 0)), VIADDMNMX (E = max(E - e, G(above))), VIMNMX3 (H = max(D, F, E))
//     FMA pipe : VIADD.16x2 (D = G(diagonal) + s + oe), VIADD.16x2 (G = H - oe)
// plus half a VIMNMX3 for the running maximum, and no register moves (the diagonal input of row r is the OLD G of row r - 1: renaming).
// The maximum is taken over the DIAGONAL candidates D, not over H: a positive maximum of H is always attained on a cell with H = D
// (a cell that takes its value from E or F is matched or beaten by the cell its gap started from), and the second use of D keeps
// ptxas from folding the add into the H maximum (VIADDMNMX + VIMNMX on the ALU pipe instead of VIADD.16x2 + VIMNMX3).  No masks
// are needed: phantom and pad columns score <= 0 on the diagonal (selector 0x88 / 0xCC: the sign byte of a table entry) and pad rows
// score the mismatch, so a cell outside the true matrix never exceeds the cell it was derived from; F is carried clamped at 0
// (F values <= 0 are interchangeable under the clamp of H), which is also what clamps H.
__device__ __forceinline__ void nvb_local_sweep4(uint32_t* __restrict__ GE, const uint8_t* __restrict__ V, int nsteps,
                                                 const uint32_t (&rowA)[8], const uint32_t (&rowB)[8],
                                                 const uint32_t (&rowC)[8], const uint32_t (&rowD)[8],
                                                 uint32_t neg_oe2, uint32_t neg_ext2, uint32_t& bestA, uint32_t& bestB) {
  uint32_t fA[8], glA[8], fB[8], glB[8];
#pragma unroll
  for (int r = 0; r < 8; r++) { fA[r] = 0; glA[r] = neg_oe2; fB[r] = 0; glB[r] = neg_oe2; }   // H = 0 left of column 0
  uint32_t g8A = neg_oe2, eoA = 0, g8B = neg_oe2, eoB = 0;      // bottom-row (G, E) of lanes A.L / A.H and B.L / B.H, handed on one step later
  uint32_t dg0A = neg_oe2, dg0B = neg_oe2;                      // G above row 0 in the previous column: row 0's diagonal input
  uint32_t prevsh = 0xCC00u;
  uint32_t selB = 0xCCCCu, selA_prev = 0xCCCCu;                 // phantom columns: score <= 0
  uint32_t ge = GE[kFastThreads];
  uint32_t v = V[kFastThreads];
  auto step = [&](int k) {
    const uint32_t selA = v | prevsh;
    prevsh = v * 256u + 0x4400u;
    const uint32_t gupA = prmt(ge, g8A, 0x5410), gupB = prmt(g8A, g8B, 0x5432);     // G of the row above each lane's strip, this column
    uint32_t eA = prmt(ge, eoA, 0x5432);
    uint32_t eB = prmt(eoA, eoB, 0x5432);
    ge = GE[(k + 2) * kFastThreads];
    v = V[(k + 2) * kFastThreads];
    uint32_t dgA = dg0A, dgB = dg0B, gA = gupA, gB = gupB, hpA = 0, hpB = 0;
    dg0A = gupA; dg0B = gupB;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      {
        const uint32_t s2 = prmt(rowA[r], rowB[r], selA);
        const uint32_t d = __vadd2(dgA, s2);
        dgA = glA[r];                                               // old G(left) of this row = the next row's diagonal input
        fA[r] = __viaddmax_s16x2_relu(fA[r], neg_ext2, glA[r]);     // max(F - e, H(left) - oe, 0)
        eA = __viaddmax_s16x2(eA, neg_ext2, gA);                    // max(E - e, H(above) - oe)
        const uint32_t h = __vimax3_s16x2(d, fA[r], eA);            // F >= 0 carries the clamp
        gA = __vadd2(h, neg_oe2); glA[r] = gA;
        if (r & 1) bestA = __vimax3_s16x2(bestA, hpA, d); else hpA = d;
      }
      {
        const uint32_t s2 = prmt(rowC[r], rowD[r], selB);
        const uint32_t d = __vadd2(dgB, s2);
        dgB = glB[r];
        fB[r] = __viaddmax_s16x2_relu(fB[r], neg_ext2, glB[r]);
        eB = __viaddmax_s16x2(eB, neg_ext2, gB);
        const uint32_t h = __vimax3_s16x2(d, fB[r], eB);
        gB = __vadd2(h, neg_oe2); glB[r] = gB;
        if (r & 1) bestB = __vimax3_s16x2(bestB, hpB, d); else hpB = d;
      }
    }
    GE[(k - 2) * kFastThreads] = prmt(gB, eB, 0x7632);          // B.H finished column k - 3 (slot k - 2): park (G, E) for the next quad-strip
    g8A = gA; eoA = eA; g8B = gB; eoB = eB;
    selB = selA_prev; selA_prev = selA;
  };
  int k = 0;
#pragma unroll 1
  for (; k + 1 < nsteps; k += 2) { step(k); step(k + 1); }
  if (k < nsteps) step(k);
}

template <int TYPE, bool SINK>
__global__ void GB200_FAST_LB k_nvb_strip(const NvbStripArgs a) {
  extern __shared__ uint32_t smem[];
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.count) return;
  const uint32_t job = a.first + idx;
  const uint32_t pi = a.pair_p ? a.pair_p[job] : job;
  const uint32_t ti = a.text_shared ? 0u : (a.pair_t ? a.pair_t[job] : job);
  if (a.pflag[pi] | a.tflag[ti]) { a.worklist[atomicAdd(a.work_count, 1u)] = job; return; }      // a symbol outside ACGT: int32 kernel
  const uint32_t po = a.poffs[pi], to = a.toffs[ti];
  const int M = (int)(a.poffs[pi + 1] - po), N = (int)(a.toffs[ti + 1] - to);
  const uint32_t* __restrict__ pw = a.ppk + (po >> 4) + pi;                // 2-bit codes, 16 per word (k_nvb_pack)
  const uint32_t* __restrict__ tw = a.tpk + (to >> 4) + ti;
  uint32_t* HE = smem + 2 * kFastThreads + threadIdx.x;                    // slot s holds column s - 1; slots -2..0 are dummies
  uint8_t* V = (uint8_t*)(smem + (a.m8max + kSlotPad) * kFastThreads) + threadIdx.x;
  const int QR = (M + 7) >> 3, ncol = QR * 8, nwT = (N + 15) >> 4;
  const int oe = a.oe, ext = a.ext, o = oe - ext;
  constexpr bool LOC = (TYPE == 1);                                        // NVBX_LOCAL
  const int bias2 = LOC ? 0 : 2 * ext;                                     // biased frame: a diagonal step carries + 2 ext
  const int tsh = LOC ? oe : bias2;                                        // LOCAL: the tables hold s + oe (diagonal inputs travel as H - oe)
  const uint32_t mis4 = ((uint32_t)(-a.mismatch + tsh) & 0xFFu) * 0x01010101u;
  const uint32_t delta = ((uint32_t)(a.match + tsh) ^ (uint32_t)(-a.mismatch + tsh)) & 0xFFu;

  // selector bytes of the pattern columns + the top boundary row
  V[0] = 0x88; HE[0] = LOC ? pack2(-oe, 0) : pack2(0, kNegSafe);      // LOCAL parks (G = H - oe, E), see nvb_local_sweep4
  for (int w = 0; w < (ncol + 15) >> 4; w++) {
    const uint32_t codes = pw[min(w, (M - 1) >> 4)];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int c = 16 * w + k;
      if (c < ncol) {
        V[(c + 1) * kFastThreads] = (c < M) ? (uint8_t)(0x80u + 0x11u * ((codes >> (2 * k)) & 3u)) : (uint8_t)0x88;
        if (LOC) HE[(c + 1) * kFastThreads] = pack2(-oe, 0);
        else {
          // H(0, j = c + 1) = Go + Ge c; biased frame of nc_sweep4: H(-1, c) carries (c - 1) ext; the parked E is the first row's own
          // E(0, c) = H(-1, c) - oe, carries c ext, and is kept as E + open
          const int htop = -(oe + ext * c);
          HE[(c + 1) * kFastThreads] = pack2(htop + (c - 1) * ext, (htop - oe) + c * ext + o);
        }
      }
    }
  }
#pragma unroll
  for (int x = 1; x <= 5; x++) { V[(ncol + x) * kFastThreads] = 0x88; HE[(ncol + x) * kFastThreads] = LOC ? pack2(-oe, 0) : pack2(0, kNegSafe); }

  const int NQ = (N + 31) >> 5;
  if (LOC) {
    const uint32_t neg_oe2 = splat16(-oe), neg_ext2 = splat16(-ext);
    uint32_t bestA = 0, bestB = 0;
    const int nsteps = M + 2 + (int)a.one;                                   // lane B.H parks column M - 1 at step M + 2; pad columns are never needed
    for (int j = 0; j < NQ; j++) {
      uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
      const uint32_t w0 = tw[min(2 * j, nwT - 1)], w1 = tw[min(2 * j + 1, nwT - 1)];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int t = 32 * j + r;
        rowA[r] = nvb_row_table(w0, 2 * r, t < N, mis4, delta);      rowB[r] = nvb_row_table(w0, 16 + 2 * r, t + 8 < N, mis4, delta);
        rowC[r] = nvb_row_table(w1, 2 * r, t + 16 < N, mis4, delta); rowD[r] = nvb_row_table(w1, 16 + 2 * r, t + 24 < N, mis4, delta);
      }
      nvb_local_sweep4(HE, V, nsteps, rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, bestA, bestB);
    }
    const uint32_t b2 = __vmaxs2(bestA, bestB);
    a.score[job] = max(lane_lo(b2), lane_hi(b2));
    return;
  }

  const uint32_t neg_o2 = splat16(-o), neg_ext2 = splat16(-ext);           // gap OPEN only: extension is free in the biased frame
  const int nsteps = ncol + 2 + (int)a.one;
  NcTrack tr; tr.best = INT_MIN / 2; tr.y = -1; tr.stop = 0; tr.cap = 0; tr.tie = 1; tr.capture = (TYPE == 0) ? 1 : 0;
  const int last_strip = (N - 1) >> 3, caprow = (N - 1) & 7;
  for (int j = 0; j < NQ; j++) {
    uint32_t rowA[8], rowB[8], rowC[8], rowD[8], pinitA[8], pinitB[8], hpA[8], hpB[8];
    const uint32_t w0 = tw[min(2 * j, nwT - 1)], w1 = tw[min(2 * j + 1, nwT - 1)];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int t = 32 * j + r;
      rowA[r] = nvb_row_table(w0, 2 * r, t < N, mis4, delta);      rowB[r] = nvb_row_table(w0, 16 + 2 * r, t + 8 < N, mis4, delta);
      rowC[r] = nvb_row_table(w1, 2 * r, t + 16 < N, mis4, delta); rowD[r] = nvb_row_table(w1, 16 + 2 * r, t + 24 < N, mis4, delta);
      if (TYPE == 0) {
        // GLOBAL: H(i, 0) = Go + Ge (i - 1) = -(o + ext i).  The diagonal input of strip row r (0-based text row t) in column 0 is
        // H(t, 0) in the frame of (row r - 1, column -1) = (r - 2) ext; the left neighbour is H(t + 1, 0) in the frame (r - 1) ext.
        const int fd = (r - 2) * ext, fh = (r - 1) * ext;
        pinitA[r] = pack2((t == 0 ? 0 : -(o + ext * t)) + fd, -(o + ext * (t + 8)) + fd);
        pinitB[r] = pack2(-(o + ext * (t + 16)) + fd, -(o + ext * (t + 24)) + fd);
        hpA[r] = pack2(-(o + ext * (t + 1)) + fh, -(o + ext * (t + 9)) + fh);
        hpB[r] = pack2(-(o + ext * (t + 17)) + fh, -(o + ext * (t + 25)) + fh);
      } else {
        pinitA[r] = pinitB[r] = splat16((r - 2) * ext);                                       // SEMI_GLOBAL: free left column, H(i, 0) = 0
      }
    }
    if (TYPE == 0) {
      const bool has_last = (last_strip >> 2) == j;
      nc_sweep4<ALGO_SEMI_GLOBAL, false>(HE, V, nsteps, rowA, rowB, rowC, rowD, pinitA, pinitB, neg_o2, neg_ext2,
                                         has_last ? M - 1 : -1000, 32 * j, N, last_strip & 3, caprow, tr, 0, 0, false, hpA, hpB);
    } else {
      nc_sweep4<ALGO_SEMI_GLOBAL, false>(HE, V, nsteps, rowA, rowB, rowC, rowD, pinitA, pinitB, neg_o2, neg_ext2, M - 1, 32 * j, N, 0, 0, tr);
    }
  }
  a.score[job] = (TYPE == 0) ? tr.cap : tr.best;
  if (SINK) { a.sink_i[job] = (TYPE == 0) ? N : tr.y + 1; a.sink_j[job] = M; }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
bool nvb_strip_eligible(int type, int match, int mismatch_pen, int oe, int ext, uint32_t max_m, uint32_t max_n, int want_sinks) {
  const uint32_t m8 = (max_m + 7u) & ~7u, n8 = (max_n + 7u) & ~7u;
  if (m8 == 0 || m8 > kFastMaxQ8 || n8 == 0) return false;
  if (type == 1 && want_sinks) return false;                      // LOCAL sinks: the reference's band-major tie order (k_nvb_s16 / k_nvb_i32w)
  if (match < 0 || match > 127 || mismatch_pen < 0 || mismatch_pen > 127) return false;
  if (ext < 0 || oe < ext || oe > 2000) return false;             // Go <= Ge <= 0
  const int64_t mn = std::min(m8, n8);
  if (type == 1) return match + oe <= 127 && (uint64_t)mn * (uint64_t)match <= 30000;  // table bytes hold s + oe; plain scores in signed 16-bit lanes
  // biased frame: the same bounds as the GASAL2 semi-global / global fast path (fast_path_eligible), with nvbio's boundaries
  if (match + 2 * ext > 127) return false;
  const int64_t span = (type == 0) ? (int64_t)m8 + n8 + 2 : (int64_t)m8 + 2;
  if (2 * (int64_t)oe + (int64_t)ext * span + mismatch_pen + 64 > 12000) return false;
  return (int64_t)(type == 0 ? mn : m8) * match + (int64_t)(m8 + 24) * ext <= 30000;
}

int nvb_strip_launch(const NvbStripArgs& a, int type, int want_sinks, cudaStream_t str) {
  const size_t smem = (size_t)(a.m8max + kSlotPad) * kFastThreads * 4 + (size_t)(a.m8max + kSlotPad) * kFastThreads;
  static bool attr_done[64] = {};
  int dev = 0; cudaGetDevice(&dev);
  bool& attr_set = attr_done[dev & 63];
  if (!attr_set) {
    const int lim = 227 * 1024;
    cudaFuncSetAttribute(k_nvb_strip<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_nvb_strip<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_nvb_strip<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_nvb_strip<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    cudaFuncSetAttribute(k_nvb_strip<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    attr_set = true;
  }
  const uint32_t g = (a.count + kFastThreads - 1) / kFastThreads;
  if (type == 0) { if (want_sinks) k_nvb_strip<0, true><<<g, kFastThreads, smem, str>>>(a); else k_nvb_strip<0, false><<<g, kFastThreads, smem, str>>>(a); }
  else if (type == 1) k_nvb_strip<1, false><<<g, kFastThreads, smem, str>>>(a);
  else if (type == 2) { if (want_sinks) k_nvb_strip<2, true><<<g, kFastThreads, smem, str>>>(a); else k_nvb_strip<2, false><<<g, kFastThreads, smem, str>>>(a); }
  else return -1;
  return 1;
}

}  // namespace gb200
