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
//   LOCAL                : banded_sweep4 with every strip's window = the whole pattern (clamped at 0, chain formulation); the score
//                          is the maximum over the true cells (pad columns are masked out by the windows, pad rows score the
//                          mismatch against everything and cannot exceed the cell above them)
// LOCAL sinks keep the reference's band-major tie order (k_nvb_s16 / k_nvb_i32w); this kernel serves scores, and the sinks of
// GLOBAL / SEMI_GLOBAL.  Jobs with a symbol outside A/C/G/T go to the int32 worklist exactly as in k_nvb_s16.
#pragma once
#include <climits>
#include "dp_s16.cuh"

namespace gb200 {

__device__ __forceinline__ uint32_t nvb_row_table(uint32_t w, int sh, bool inside, uint32_t mis4, uint32_t delta) {
  return inside ? (mis4 ^ (delta << (8u * ((w >> sh) & 3u)))) : mis4;
}

template <int TYPE, bool SINK>
__global__ void __launch_bounds__(kFastThreads) k_nvb_strip(const NvbStripArgs a) {
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
  const uint32_t mis4 = ((uint32_t)(-a.mismatch + bias2) & 0xFFu) * 0x01010101u;
  const uint32_t delta = ((uint32_t)(a.match + bias2) ^ (uint32_t)(-a.mismatch + bias2)) & 0xFFu;

  // selector bytes of the pattern columns + the top boundary row
  V[0] = 0x88; HE[0] = LOC ? 0u : pack2(0, kNegSafe);
  for (int w = 0; w < (ncol + 15) >> 4; w++) {
    const uint32_t codes = pw[min(w, (M - 1) >> 4)];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int c = 16 * w + k;
      if (c < ncol) {
        V[(c + 1) * kFastThreads] = (c < M) ? (uint8_t)(0x80u + 0x11u * ((codes >> (2 * k)) & 3u)) : (uint8_t)0x88;
        if (LOC) HE[(c + 1) * kFastThreads] = 0u;
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
  for (int x = 1; x <= 5; x++) { V[(ncol + x) * kFastThreads] = 0x88; HE[(ncol + x) * kFastThreads] = LOC ? 0u : pack2(0, kNegSafe); }

  const int NQ = (N + 31) >> 5;
  if (LOC) {
    const uint32_t neg_oe2 = splat16(-oe), neg_ext2 = splat16(-ext);
    LaneBest4 lb; lanebest_init(lb.a); lanebest_init(lb.b);
    for (int j = 0; j < NQ; j++) {
      uint32_t rowA[8], rowB[8], rowC[8], rowD[8];
      const uint32_t w0 = tw[min(2 * j, nwT - 1)], w1 = tw[min(2 * j + 1, nwT - 1)];
      BandWin bw;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const bool have = 32 * j + 8 * i < N;
        bw.enter[i] = have ? i : 0; bw.exitk[i] = have ? M + i : 0; bw.tile_base[i] = 0;     // the window of a strip: columns [0, M)
      }
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int t = 32 * j + r;
        rowA[r] = nvb_row_table(w0, 2 * r, t < N, mis4, delta);      rowB[r] = nvb_row_table(w0, 16 + 2 * r, t + 8 < N, mis4, delta);
        rowC[r] = nvb_row_table(w1, 2 * r, t + 16 < N, mis4, delta); rowD[r] = nvb_row_table(w1, 16 + 2 * r, t + 24 < N, mis4, delta);
      }
      banded_sweep4<false>(HE, V, bw, rowA, rowB, rowC, rowD, neg_oe2, neg_ext2, a.eight, a.one, (uint32_t)j << 16, lb);
    }
    const Winner w_ = pick_winner4(lb);
    a.score[job] = w_.score;
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
  if (type == 1) return (uint64_t)mn * (uint64_t)match <= 4000;   // keys are score * 8 + 7 in a signed 16-bit lane
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
