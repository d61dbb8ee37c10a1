// nw_matrix.cu -- full-matrix linear-gap Needleman-Wunsch in ONE persistent launch (C ABI: include/nw_b200.h).
//
// What the reference does (Non-CDP/NW): nw_cpu (main.cu:9-27) fills matrix[q*N + r]; nw_gpu3 (kernel3.cu:8-103) walks the
// anti-diagonals of 64 x 64 tiles with one launch + cudaDeviceSynchronize per tile diagonal, every tile re-reading its
// halo from the matrix in HBM.  The kernel is bound by the 4 B per cell it must WRITE (N = 32000: 4.1 GB), so the design
// point here is the HBM write roofline, not the integer pipes.
//
// This kernel: a warp owns a ROW-BLOCK of 128 query rows (lane L: rows 4L..4L+3) and walks the reference 4 columns per
// step, lanes skewed by one step, so lane L receives the bottom row of lane L-1's previous 4 x 4 tile by __shfl_up.  Lane
// 31's bottom row is the next row-block's top boundary: it goes through a double-buffered array of self-validating
// 8-byte (value, tag) entries in L2 (no flags, no fences; the consumer re-reads a chunk until every tag names this launch
// and the producing block -- the protocol of dp_wave.cuh) and is prefetched one chunk ahead.  Row-blocks are handed out by
// a ticket counter in row order, so a block only ever waits for a lower ticket that is resident or finished: deadlock-free
// for any grid.  Stores are staged through shared memory so that every STG.128 of a warp writes four 128-byte row segments
// (a lane's 4 x 4 tile alone would be 32 scattered 16-byte sector fragments per store).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include "../../include/nw_b200.h"

namespace nwb200 {

#ifndef NW_ROWS_PER_LANE
#define NW_ROWS_PER_LANE 4
#endif
#ifndef NW_UNROLL_STEPS
#define NW_UNROLL_STEPS 8                  // the step body must stay inside the instruction cache: one warp per scheduler cannot hide fetch misses
#endif
#ifndef NW_CHUNK_STEPS
#define NW_CHUNK_STEPS 8
#endif
#ifndef NW_PREFETCH_STEPS
#define NW_PREFETCH_STEPS 5
#endif
constexpr int kR = NW_ROWS_PER_LANE;       // rows per lane (1, 2 or 4): the step time scales with kR, the number of lane-skew steps with 1 / kR
constexpr int kC = 4;                      // columns per step
#ifndef NW_FLUSH_UNROLL
#define NW_FLUSH_UNROLL 8
#endif
constexpr int kFlushBatch = NW_FLUSH_UNROLL;
#define NW_PRAGMA(x) _Pragma(#x)
#define NW_UNROLL(n) NW_PRAGMA(unroll n)      // nvcc ignores a constexpr unroll factor: it has to be a literal
constexpr int kBlockRows = 32 * kR;        // rows per row-block (one warp)
constexpr int kGroup = 8;                  // steps per staged row segment: 8 x 4 columns x 4 B = 128 B
constexpr int kChunk = NW_CHUNK_STEPS;     // steps per boundary chunk
constexpr int kPrefetch = NW_PREFETCH_STEPS;   // a chunk is requested this many steps before its first use
constexpr int kStride = kR * kGroup * kC + 4;
constexpr uint32_t kFull = 0xFFFFFFFFu;
static_assert((32 * kR / 4) % kFlushBatch == 0, "flush batches");
static_assert(kGroup % kChunk == 0 && kPrefetch < kChunk && (kR == 1 || kR == 2 || kR == 4 || kR == 8) && kC == 4, "flush / chunk schedule assumptions");

struct NwArgs {
  const uint32_t* ref32;                   // reference, 4 bytes per word, zero-padded to Spad words (+ slack)
  const uint8_t* qry;
  int* mat;
  uint32_t N;
  int match, mismatch, ins, del;
  uint2* bnd;                              // [2][Spad * 4] (value, tag)
  uint32_t* ticket;
  uint32_t seq, nblocks, Spad;
  int one;                                 // == 1 at run time: keeps `up * one + ins` an IMAD (FMA pipe) instead of an ALU-pipe add
  unsigned long long* dbg;                 // optional [nblocks][4]: start ns, end ns, failed chunk polls, SM id (NWX_DEBUG_TIMING=1)
};

__global__ void k_nw_prep(const uint8_t* __restrict__ ref, uint32_t N, uint32_t* __restrict__ ref32, uint32_t words, uint32_t* ticket) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w == 0) *ticket = 0;
  if (w >= words) return;
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { const uint32_t c = 4 * w + k; if (c < N) v |= (uint32_t)ref[c] << (8 * k); }
  ref32[w] = v;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(32) k_nw_matrix(const NwArgs a) {
  __shared__ __align__(16) int stage[32 * kStride];                  // lane region: [row 0..3][8 steps x 4 columns] (+16 B pad: conflict-free STS.128)
  __shared__ __align__(16) int4 bndbuf[2][kChunk];
  const int lane = threadIdx.x;
  const uint32_t N = a.N;
  const int Spad = (int)a.Spad, nsteps = Spad + 32;
  const int match = a.match, mismatch = a.mismatch, ins = a.ins, del = a.del;

  for (;;) {
    uint32_t blk = 0;
    if (lane == 0) blk = atomicAdd(a.ticket, 1u);
    blk = __shfl_sync(kFull, blk, 0);
    if (blk >= a.nblocks) return;

    unsigned long long t_start = 0; uint32_t n_spin = 0;
    if (a.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    const uint32_t q0 = blk * kBlockRows + lane * kR;
    int qb[kR], left[kR], bot[kC];
#pragma unroll
    for (int i = 0; i < kR; i++) {
      qb[i] = (q0 + i < N) ? (int)a.qry[q0 + i] : 256;              // rows past N never match, never stored
      left[i] = (int)(q0 + i + 1) * ins;                            // H(q, -1) = (q + 1) * INSERTION   (main.cu:14)
    }
#pragma unroll
    for (int j = 0; j < kC; j++) bot[j] = 0;
    int corner = (int)q0 * ins;                                     // H(q0 - 1, -1); 0 for the first row (main.cu:15)

    const uint32_t tag_in = (a.seq << 12) | ((blk - 1) & 0xFFFu), tag_out = (a.seq << 12) | (blk & 0xFFFu);
    const uint2* bnd_in = a.bnd + (size_t)((blk - 1) & 1u) * (size_t)Spad * kC;
    uint2* bnd_out = a.bnd + (size_t)(blk & 1u) * (size_t)Spad * kC;
    const bool produce = blk + 1 < a.nblocks, consume = blk > 0;

    // ---- boundary chunks: lanes t < kChunk hold step (chunk * kChunk + t) of the chunk being fetched ----
    uint4 pfA = make_uint4(0, 0, 0, 0), pfB = make_uint4(0, 0, 0, 0);
    auto chunk_request = [&](int chunk) {
      const int ck = chunk * kChunk + lane;
      if (lane < kChunk && ck < Spad) {
        const uint4* p = (const uint4*)(bnd_in + (size_t)ck * kC);
        pfA = __ldcg(p); pfB = __ldcg(p + 1);
      }
    };
    auto chunk_commit = [&](int chunk) {                            // spin until every entry carries this launch's tag
      const int ck = chunk * kChunk + lane;
      for (;;) {
        const bool ok = lane >= kChunk || ck >= Spad || (pfA.y == tag_in && pfA.w == tag_in && pfB.y == tag_in && pfB.w == tag_in);
        if (__all_sync(kFull, ok)) break;
        n_spin++;
        if (!ok) { const uint4* p = (const uint4*)(bnd_in + (size_t)ck * kC); pfA = __ldcg(p); pfB = __ldcg(p + 1); }
      }
      if (lane < kChunk) bndbuf[chunk & 1][lane] = make_int4((int)pfA.x, (int)pfA.z, (int)pfB.x, (int)pfB.z);
      __syncwarp();
    };
    if (consume) { chunk_request(0); chunk_commit(0); }

    uint32_t rw_next = a.ref32[0];                                  // lanes > 0 start before column 0: clamped
    for (int s0 = 0; s0 < nsteps; s0 += kGroup) {
NW_UNROLL(NW_UNROLL_STEPS)
      for (int u = 0; u < kGroup; u++) {
        const int s = s0 + u;
        const int cs = s - lane;                                    // this lane's column step
        const bool active = cs >= 0 && cs < Spad;
        if (consume) {
          if (u % kChunk == 0 && s > 0 && s < Spad) chunk_commit(s / kChunk);
          if (u % kChunk == kChunk - kPrefetch && (s / kChunk + 1) * kChunk < Spad) chunk_request(s / kChunk + 1);
        }
        // top row of this step's tile: bottom row of lane - 1's previous tile / the row-block above / the virtual row -1
        int4 bv = bndbuf[(s / kChunk) & 1][u % kChunk];
        if (!consume) { const int c = kC * s; bv = make_int4((c + 1) * del, (c + 2) * del, (c + 3) * del, (c + 4) * del); }   // H(-1, r) = (r + 1) * DELETION (main.cu:13)
        int t0 = __shfl_up_sync(kFull, bot[0], 1), t1 = __shfl_up_sync(kFull, bot[1], 1);
        int t2 = __shfl_up_sync(kFull, bot[2], 1), t3 = __shfl_up_sync(kFull, bot[3], 1);
        if (lane == 0) { t0 = bv.x; t1 = bv.y; t2 = bv.z; t3 = bv.w; }
        const uint32_t rw = rw_next;
        { const int cn = cs + 1; rw_next = a.ref32[cn < 0 ? 0 : cn]; }
        if (active) {
          const int top[kC] = {t0, t1, t2, t3};
          int rb[kC];
#pragma unroll
          for (int j = 0; j < kC; j++) rb[j] = (int)((rw >> (8 * j)) & 0xFFu);
          int h[kR][kC];
#pragma unroll
          for (int i = 0; i < kR; i++) {
#pragma unroll
            for (int j = 0; j < kC; j++) {
              const int tl = (i == 0) ? (j == 0 ? corner : top[j - 1]) : (j == 0 ? left[i - 1] : h[i - 1][j - 1]);
              const int up = (i == 0) ? top[j] : h[i - 1][j];
              const int lf = (j == 0) ? left[i] : h[i][j - 1];
              h[i][j] = __vimax3_s32(tl + (qb[i] == rb[j] ? match : mismatch), up + ins, lf + del);   // main.cu:17-23
            }
          }
          corner = t3;
#pragma unroll
          for (int i = 0; i < kR; i++) {
            left[i] = h[i][kC - 1];
            *(int4*)&stage[lane * kStride + i * (kGroup * kC) + u * kC] = make_int4(h[i][0], h[i][1], h[i][2], h[i][3]);
          }
#pragma unroll
          for (int j = 0; j < kC; j++) bot[j] = h[kR - 1][j];
          if (lane == 31 && produce) {
            uint4* p = (uint4*)(bnd_out + (size_t)cs * kC);
            __stcg(p, make_uint4((uint32_t)bot[0], tag_out, (uint32_t)bot[1], tag_out));
            __stcg(p + 1, make_uint4((uint32_t)bot[2], tag_out, (uint32_t)bot[3], tag_out));
          }
        }
      }
      __syncwarp();
      // flush: every lane's region now holds its last 8 steps = 4 rows x 32 consecutive columns (lane m: columns
      // 4 (s0 - m) ...).  One LDS.128 + STG.128 per region, 8 lanes per 128-byte row segment.
#pragma unroll 1
      for (int it0 = 0; it0 < 32 * kR / 4; it0 += kFlushBatch) {   // 4 block rows per LDS.128; all loads of a batch are issued before its stores
        int4 v[kFlushBatch];
#pragma unroll
        for (int k = 0; k < kFlushBatch; k++) {
          const int brow = 4 * (it0 + k) + (lane >> 3), m = brow / kR, i = brow % kR;
          v[k] = *(const int4*)&stage[m * kStride + i * (kGroup * kC) + (lane & 7) * 4];
        }
#pragma unroll
        for (int k = 0; k < kFlushBatch; k++) {
          const int brow = 4 * (it0 + k) + (lane >> 3), m = brow / kR;
          const int csb = s0 - m + (lane & 7);
          const uint32_t q = blk * kBlockRows + brow, col = (uint32_t)(kC * csb);
          if (csb >= 0 && csb < Spad && q < N && col < N) {
            int* dst = a.mat + (size_t)q * N + col;
            if (ALIGNED) *(int4*)dst = v[k];
            else { dst[0] = v[k].x; if (col + 1 < N) dst[1] = v[k].y; if (col + 2 < N) dst[2] = v[k].z; if (col + 3 < N) dst[3] = v[k].w; }
          }
        }
      }
      __syncwarp();
    }
    if (a.dbg && lane == 0) {
      unsigned long long t_end; uint32_t smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.dbg[4 * blk] = t_start; a.dbg[4 * blk + 1] = t_end; a.dbg[4 * blk + 2] = n_spin; a.dbg[4 * blk + 3] = smid;
    }
  }
}


// ---- warp-specialised variant ---------------------------------------------------------------------------------------
// One scheduler issues at most one ALU-pipe instruction every other cycle for a single warp, and the row-block pipeline is a
// chain of single warps: whatever the compute warp does not have to do itself shortens the critical path of the whole
// matrix.  A CTA is TWO warps: warp 0 computes tiles and only stores them to a double-buffered shared-memory stage; warp 1
// (on another scheduler) turns finished 8-step groups into coalesced global stores and fetches / validates the boundary
// chunks of the row-block above.  Hand-shakes are named barriers (producer bar.arrive, consumer bar.sync, 64 threads):
//   FULL[p]  compute -> helper : stage buffer p holds a finished group
//   FREE[p]  helper -> compute : stage buffer p has been written out
//   BND[p]   helper -> compute : boundary chunk buffer p is valid
__device__ __forceinline__ void nb_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
constexpr int kBarFull = 1, kBarFree = 3, kBarBnd = 5;

template <bool ALIGNED>
__global__ void __launch_bounds__(64) k_nw_matrix_ws(const NwArgs a) {
  __shared__ __align__(16) int stage[2][32 * kStride];
  __shared__ __align__(16) int4 bndbuf[2][kGroup];
  __shared__ uint32_t s_blk;
  const int lane = threadIdx.x & 31;
  const bool helper = threadIdx.x >= 32;
  const uint32_t N = a.N;
  const int Spad = (int)a.Spad, nsteps = Spad + 32, G = nsteps / kGroup;
  const int match = a.match, mismatch = a.mismatch, ins = a.ins, del = a.del, one = a.one;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_blk = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t blk = s_blk;
    if (blk >= a.nblocks) return;
    const bool produce = blk + 1 < a.nblocks, consume = blk > 0;
    const uint32_t tag_in = (a.seq << 12) | ((blk - 1) & 0xFFFu), tag_out = (a.seq << 12) | (blk & 0xFFFu);

    if (helper) {
      const uint2* bnd_in = a.bnd + (size_t)((blk - 1) & 1u) * (size_t)Spad * kC;
      uint4 pfA = make_uint4(0, 0, 0, 0), pfB = make_uint4(0, 0, 0, 0);
      auto request = [&](int chunk) {
        const int ck = chunk * kGroup + lane;
        if (lane < kGroup && ck < Spad) { const uint4* p = (const uint4*)(bnd_in + (size_t)ck * kC); pfA = __ldcg(p); pfB = __ldcg(p + 1); }
      };
      auto commit = [&](int chunk) {
        const int ck = chunk * kGroup + lane;
        for (;;) {
          const bool ok = lane >= kGroup || ck >= Spad || (pfA.y == tag_in && pfA.w == tag_in && pfB.y == tag_in && pfB.w == tag_in);
          if (__all_sync(kFull, ok)) break;
          if (!ok) { const uint4* p = (const uint4*)(bnd_in + (size_t)ck * kC); pfA = __ldcg(p); pfB = __ldcg(p + 1); }
        }
        if (lane < kGroup) bndbuf[chunk & 1][lane] = make_int4((int)pfA.x, (int)pfA.z, (int)pfB.x, (int)pfB.z);
        nb_arrive(kBarBnd + (chunk & 1));
      };
      if (consume) { request(0); commit(0); if (kGroup < Spad) request(1); }
      for (int g = 0; g < G; g++) {
        const int p = g & 1, s0 = g * kGroup;
        if (consume && (g + 1) * kGroup < Spad) { commit(g + 1); if ((g + 2) * kGroup < Spad) request(g + 2); }
        nb_sync(kBarFull + p);
        const int* st = stage[p];
#pragma unroll 1
        for (int it0 = 0; it0 < 32 * kR / 4; it0 += kFlushBatch) {
          int4 v[kFlushBatch];
#pragma unroll
          for (int k = 0; k < kFlushBatch; k++) {
            const int brow = 4 * (it0 + k) + (lane >> 3), m = brow / kR, i = brow % kR;
            v[k] = *(const int4*)&st[m * kStride + i * (kGroup * kC) + (lane & 7) * 4];
          }
#pragma unroll
          for (int k = 0; k < kFlushBatch; k++) {
            const int brow = 4 * (it0 + k) + (lane >> 3), m = brow / kR;
            const int csb = s0 - m + (lane & 7);
            const uint32_t q = blk * kBlockRows + brow, col = (uint32_t)(kC * csb);
            if (csb >= 0 && csb < Spad && q < N && col < N) {
              int* dst = a.mat + (size_t)q * N + col;
              if (ALIGNED) *(int4*)dst = v[k];
              else { dst[0] = v[k].x; if (col + 1 < N) dst[1] = v[k].y; if (col + 2 < N) dst[2] = v[k].z; if (col + 3 < N) dst[3] = v[k].w; }
            }
          }
        }
        if (g + 2 < G) nb_arrive(kBarFree + p);
      }
      continue;
    }

    // ---- compute warp ----
    unsigned long long t_start = 0;
    if (a.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    const uint32_t q0 = blk * kBlockRows + lane * kR;
    int qb[kR], left[kR], bot[kC];
#pragma unroll
    for (int i = 0; i < kR; i++) {
      qb[i] = (q0 + i < N) ? (int)a.qry[q0 + i] : 256;
      left[i] = (int)(q0 + i + 1) * ins;
    }
#pragma unroll
    for (int j = 0; j < kC; j++) bot[j] = 0;
    int corner = (int)q0 * ins;
    uint2* bnd_out = a.bnd + (size_t)(blk & 1u) * (size_t)Spad * kC;
    uint32_t rw_next = a.ref32[0];
    for (int g = 0; g < G; g++) {
      const int p = g & 1, s0 = g * kGroup;
      if (g >= 2) nb_sync(kBarFree + p);
      if (consume && s0 < Spad) nb_sync(kBarBnd + p);
      int* st = stage[p] + lane * kStride;
NW_UNROLL(NW_UNROLL_STEPS)
      for (int u = 0; u < kGroup; u++) {
        const int s = s0 + u;
        const int cs = s - lane;
        const bool active = cs >= 0 && cs < Spad;
        int4 bv = bndbuf[p][u];
        if (!consume) { const int c = kC * s; bv = make_int4((c + 1) * del, (c + 2) * del, (c + 3) * del, (c + 4) * del); }
        int t0 = __shfl_up_sync(kFull, bot[0], 1), t1 = __shfl_up_sync(kFull, bot[1], 1);
        int t2 = __shfl_up_sync(kFull, bot[2], 1), t3 = __shfl_up_sync(kFull, bot[3], 1);
        if (lane == 0) { t0 = bv.x; t1 = bv.y; t2 = bv.z; t3 = bv.w; }
        const uint32_t rw = rw_next;
        { const int cn = cs + 1; rw_next = a.ref32[cn < 0 ? 0 : cn]; }
        if (active) {
          const int top[kC] = {t0, t1, t2, t3};
          int rb[kC];
#pragma unroll
          for (int j = 0; j < kC; j++) rb[j] = (int)((rw >> (8 * j)) & 0xFFu);
          int h[kR][kC];
#pragma unroll
          for (int i = 0; i < kR; i++) {
#pragma unroll
            for (int j = 0; j < kC; j++) {
              const int tl = (i == 0) ? (j == 0 ? corner : top[j - 1]) : (j == 0 ? left[i - 1] : h[i - 1][j - 1]);
              const int up = (i == 0) ? top[j] : h[i - 1][j];
              const int lf = (j == 0) ? left[i] : h[i][j - 1];
              h[i][j] = __vimax3_s32(tl + (qb[i] == rb[j] ? match : mismatch), up * one + ins, lf + del);
            }
          }
          corner = t3;
#pragma unroll
          for (int i = 0; i < kR; i++) {
            left[i] = h[i][kC - 1];
            *(int4*)&st[i * (kGroup * kC) + u * kC] = make_int4(h[i][0], h[i][1], h[i][2], h[i][3]);
          }
#pragma unroll
          for (int j = 0; j < kC; j++) bot[j] = h[kR - 1][j];
          if (lane == 31 && produce) {
            uint4* pp = (uint4*)(bnd_out + (size_t)cs * kC);
            __stcg(pp, make_uint4((uint32_t)bot[0], tag_out, (uint32_t)bot[1], tag_out));
            __stcg(pp + 1, make_uint4((uint32_t)bot[2], tag_out, (uint32_t)bot[3], tag_out));
          }
        }
      }
      nb_arrive(kBarFull + p);
    }
    if (a.dbg && lane == 0) {
      unsigned long long t_end; uint32_t smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.dbg[4 * blk] = t_start; a.dbg[4 * blk + 1] = t_end; a.dbg[4 * blk + 2] = 0; a.dbg[4 * blk + 3] = smid;
    }
  }
}

__global__ void __launch_bounds__(256) k_nw_checksum(const int* __restrict__ m, uint64_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    acc += (unsigned long long)(uint32_t)m[i] * (unsigned long long)(((uint32_t)i * 2654435761u) | 1u);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(kFull, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ---- host side --------------------------------------------------------------------------------------------------
struct Scratch {
  uint32_t* ref32 = nullptr; size_t ref_words = 0;
  uint2* bnd = nullptr; size_t bnd_entries = 0;
  uint32_t* ticket = nullptr;
  unsigned long long* sum = nullptr;
  uint32_t seq = 0;
  uint8_t* d_ref = nullptr; uint8_t* d_qry = nullptr; size_t seq_cap = 0;     // nwx_matrix_host
  int* d_mat = nullptr; size_t mat_elems = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t busy = nullptr;          // recorded after every launch: the next launch (any stream) waits on it before it reuses the scratch
  int sms = 0;
};
constexpr int kMaxDev = 64;
Scratch g_scratch[kMaxDev];
std::mutex g_mu;
uint64_t g_launches = 0;
int g_variant = 0;                        // 0: warp-specialised (compute + store/fetch helper warp), 1: single warp per row-block

#define NW_CHECK(x)                                                                                             \
  do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[NW CUDA ERROR:] %s (%s:%d)\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while (0)

Scratch* scratch_for_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  Scratch* sc = &g_scratch[dev];
  if (sc->sms == 0) cudaDeviceGetAttribute(&sc->sms, cudaDevAttrMultiProcessorCount, dev);
  return sc;
}

int launch(const uint8_t* ref_d, const uint8_t* qry_d, int* mat_d, uint32_t N, const nwx_scores_t* scores, cudaStream_t st) {
  if (N == 0) return 0;
  if (!ref_d || !qry_d || !mat_d) { fprintf(stderr, "[NW ERROR:] null device pointer\n"); return -1; }
  Scratch* sc = scratch_for_current_device();
  if (!sc) { fprintf(stderr, "[NW CUDA ERROR:] no CUDA device\n"); return -1; }
  const uint32_t S = (N + kC - 1) / kC, Spad = (S + kGroup - 1) / kGroup * kGroup;
  const size_t words = (size_t)Spad + 64, entries = 2 * (size_t)Spad * kC;
  // ONE scratch (ref32, boundary ring, ticket, tags) per device: launches on different streams are ordered on the device through
  // an event, so a second call cannot overwrite the ring while the first matrix kernel is still running
  if (!sc->busy) NW_CHECK(cudaEventCreateWithFlags(&sc->busy, cudaEventDisableTiming));
  else NW_CHECK(cudaStreamWaitEvent(st, sc->busy, 0));
  if (sc->ref_words < words) { if (sc->ref32) cudaFree(sc->ref32); sc->ref32 = nullptr; NW_CHECK(cudaMalloc(&sc->ref32, words * sizeof(uint32_t))); sc->ref_words = words; }
  if (sc->bnd_entries < entries) {
    if (sc->bnd) cudaFree(sc->bnd);
    sc->bnd = nullptr; NW_CHECK(cudaMalloc(&sc->bnd, entries * sizeof(uint2))); sc->bnd_entries = entries;
    NW_CHECK(cudaMemsetAsync(sc->bnd, 0, entries * sizeof(uint2), st)); sc->seq = 0;
  }
  if (!sc->ticket) NW_CHECK(cudaMalloc(&sc->ticket, sizeof(uint32_t)));
  if (++sc->seq >= (1u << 20)) { NW_CHECK(cudaMemsetAsync(sc->bnd, 0, sc->bnd_entries * sizeof(uint2), st)); sc->seq = 1; }   // tag 0 = never written
  NwArgs a;
  a.ref32 = sc->ref32; a.qry = qry_d; a.mat = mat_d; a.N = N;
  a.match = scores ? scores->match : 1; a.mismatch = scores ? scores->mismatch : -1;
  a.ins = scores ? scores->insertion : -1; a.del = scores ? scores->deletion : -1;
  a.bnd = sc->bnd; a.ticket = sc->ticket; a.seq = sc->seq; a.nblocks = (N + kBlockRows - 1) / kBlockRows; a.Spad = Spad;
  a.dbg = nullptr; a.one = 1;
  const bool debug = getenv("NWX_DEBUG_TIMING") != nullptr;
  if (debug) NW_CHECK(cudaMalloc(&a.dbg, (size_t)a.nblocks * 4 * sizeof(unsigned long long)));
  k_nw_prep<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(ref_d, N, sc->ref32, (uint32_t)words, sc->ticket);
  const uint32_t grid = a.nblocks < (uint32_t)sc->sms * 8u ? a.nblocks : (uint32_t)sc->sms * 8u;
  const bool aligned = (N % 4 == 0) && (((uintptr_t)mat_d & 15u) == 0);
  if (g_variant == 0) { if (aligned) k_nw_matrix_ws<true><<<grid, 64, 0, st>>>(a); else k_nw_matrix_ws<false><<<grid, 64, 0, st>>>(a); }
  else { if (aligned) k_nw_matrix<true><<<grid, 32, 0, st>>>(a); else k_nw_matrix<false><<<grid, 32, 0, st>>>(a); }
  NW_CHECK(cudaGetLastError());
  NW_CHECK(cudaEventRecord(sc->busy, st));
  g_launches += 2;
  if (debug) {      // development aid: per-row-block timeline on stderr (synchronises)
    unsigned long long* h = (unsigned long long*)malloc((size_t)a.nblocks * 4 * sizeof(unsigned long long));
    NW_CHECK(cudaMemcpyAsync(h, a.dbg, (size_t)a.nblocks * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    NW_CHECK(cudaStreamSynchronize(st));
    const unsigned long long t0 = h[0];
    for (uint32_t b = 0; b < a.nblocks; b += (a.nblocks > 40 ? a.nblocks / 20 : 1))
      fprintf(stderr, "[nw dbg] blk %4u start %9.1f us end %9.1f us (dur %8.1f us, %.0f ns/step) spins %llu sm %llu\n", b, (h[4 * b] - t0) * 1e-3, (h[4 * b + 1] - t0) * 1e-3,
              (h[4 * b + 1] - h[4 * b]) * 1e-3, (double)(h[4 * b + 1] - h[4 * b]) / (Spad + 32), h[4 * b + 2], h[4 * b + 3]);
    free(h); cudaFree(a.dbg);
  }
  return 0;
}

int checksum(const int* mat_d, uint64_t n, cudaStream_t st, uint64_t* out) {
  Scratch* sc = scratch_for_current_device();
  if (!sc) return -1;
  if (!sc->sum) NW_CHECK(cudaMalloc(&sc->sum, sizeof(unsigned long long)));
  NW_CHECK(cudaMemsetAsync(sc->sum, 0, sizeof(unsigned long long), st));
  if (n) { k_nw_checksum<<<sc->sms * 8, 256, 0, st>>>(mat_d, n, sc->sum); NW_CHECK(cudaGetLastError()); g_launches++; }
  unsigned long long h = 0;
  NW_CHECK(cudaMemcpyAsync(&h, sc->sum, sizeof(h), cudaMemcpyDeviceToHost, st));
  NW_CHECK(cudaStreamSynchronize(st));
  *out = h;
  return 0;
}

}  // namespace nwb200

extern "C" {

int nwx_matrix_async(const unsigned char* reference_d, const unsigned char* query_d, int* matrix_d, unsigned int N,
                     const nwx_scores_t* scores, void* cuda_stream) {
  std::lock_guard<std::mutex> lk(nwb200::g_mu);      // one scratch per device: launches on a device are serialised by the caller's stream order
  return nwb200::launch(reference_d, query_d, matrix_d, N, scores, (cudaStream_t)cuda_stream);
}

void nw_gpu_b200(unsigned char* reference_d, unsigned char* query_d, int* matrix_d, unsigned int N) {
  if (nwx_matrix_async(reference_d, query_d, matrix_d, N, nullptr, nullptr) != 0) exit(EXIT_FAILURE);
  cudaError_t e = cudaStreamSynchronize(nullptr);     // the reference's nw_gpuX return after cudaDeviceSynchronize (kernel3.cu:94,101)
  if (e != cudaSuccess) { fprintf(stderr, "[NW CUDA ERROR:] %s\n", cudaGetErrorString(e)); exit(EXIT_FAILURE); }
}

uint64_t nwx_checksum(const int* matrix_d, uint64_t n_elements, void* cuda_stream) {
  std::lock_guard<std::mutex> lk(nwb200::g_mu);
  uint64_t h = 0;
  if (nwb200::checksum(matrix_d, n_elements, (cudaStream_t)cuda_stream, &h) != 0) { fprintf(stderr, "[NW CUDA ERROR:] checksum failed\n"); exit(EXIT_FAILURE); }
  return h;
}

int nwx_matrix_host(const unsigned char* reference, const unsigned char* query, int* matrix_host, unsigned int N,
                    const nwx_scores_t* scores, float ms[3], uint64_t* checksum) {
  using namespace nwb200;
  std::lock_guard<std::mutex> lk(g_mu);
  Scratch* sc = scratch_for_current_device();
  if (!sc) { fprintf(stderr, "[NW CUDA ERROR:] no CUDA device\n"); return -1; }
  if (N == 0) { if (ms) ms[0] = ms[1] = ms[2] = 0.f; if (checksum) *checksum = 0; return 0; }
  if (!reference || !query) { fprintf(stderr, "[NW ERROR:] null host pointer\n"); return -1; }
  if (sc->seq_cap < N) {
    if (sc->d_ref) cudaFree(sc->d_ref); if (sc->d_qry) cudaFree(sc->d_qry);
    sc->d_ref = sc->d_qry = nullptr;
    NW_CHECK(cudaMalloc(&sc->d_ref, N)); NW_CHECK(cudaMalloc(&sc->d_qry, N)); sc->seq_cap = N;
  }
  const size_t elems = (size_t)N * N;
  if (sc->mat_elems < elems) { if (sc->d_mat) cudaFree(sc->d_mat); sc->d_mat = nullptr; NW_CHECK(cudaMalloc(&sc->d_mat, elems * sizeof(int))); sc->mat_elems = elems; }
  for (int e = 0; e < 4; e++) if (!sc->ev[e]) NW_CHECK(cudaEventCreate(&sc->ev[e]));
  cudaStream_t st = nullptr;
  NW_CHECK(cudaEventRecord(sc->ev[0], st));
  NW_CHECK(cudaMemcpyAsync(sc->d_ref, reference, N, cudaMemcpyHostToDevice, st));
  NW_CHECK(cudaMemcpyAsync(sc->d_qry, query, N, cudaMemcpyHostToDevice, st));
  NW_CHECK(cudaEventRecord(sc->ev[1], st));
  if (launch(sc->d_ref, sc->d_qry, sc->d_mat, N, scores, st) != 0) return -1;
  NW_CHECK(cudaEventRecord(sc->ev[2], st));
  if (matrix_host) NW_CHECK(cudaMemcpyAsync(matrix_host, sc->d_mat, elems * sizeof(int), cudaMemcpyDeviceToHost, st));
  NW_CHECK(cudaEventRecord(sc->ev[3], st));
  NW_CHECK(cudaStreamSynchronize(st));
  if (ms) for (int k = 0; k < 3; k++) NW_CHECK(cudaEventElapsedTime(&ms[k], sc->ev[k], sc->ev[k + 1]));
  if (checksum) { if (nwb200::checksum(sc->d_mat, elems, st, checksum) != 0) return -1; }
  return 0;
}

uint64_t nwx_kernel_launches(void) { return nwb200::g_launches; }
void nwx_set_variant(int variant) { nwb200::g_variant = variant; }

void nwx_release(void) {
  using namespace nwb200;
  std::lock_guard<std::mutex> lk(g_mu);
  int cur = 0; cudaGetDevice(&cur);
  for (int d = 0; d < kMaxDev; d++) {
    Scratch& sc = g_scratch[d];
    if (!sc.ref32 && !sc.bnd && !sc.ticket && !sc.sum && !sc.d_ref && !sc.d_mat && !sc.ev[0] && !sc.busy) continue;
    cudaSetDevice(d);
    cudaFree(sc.ref32); cudaFree(sc.bnd); cudaFree(sc.ticket); cudaFree(sc.sum); cudaFree(sc.d_ref); cudaFree(sc.d_qry); cudaFree(sc.d_mat);
    for (int e = 0; e < 4; e++) if (sc.ev[e]) cudaEventDestroy(sc.ev[e]);
    if (sc.busy) cudaEventDestroy(sc.busy);
    sc = Scratch();
  }
  cudaSetDevice(cur);
}

}  // extern "C"
