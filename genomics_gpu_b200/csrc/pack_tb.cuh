// pack_tb.cuh -- K1 (4-bit pack + reverse/complement) and K4 (traceback walker -> CIGAR bytes).
// Reference: Non-CDP/GASAL2/src/kernels/pack_rc_seqs.h (gasal_pack_kernel :13-53,
// gasal_reversecomplement_kernel :56-212) and kernels/get_tb.h (gasal_get_tb :4-149).
#pragma once
#include "dp_common.cuh"

namespace gb200 {

// ---- K1a: pack -------------------------------------------------------------------------------
// 16 ASCII bytes (one uint4 load) -> two packed words (one uint2 store) per thread-iteration;
// nibble = byte & 15, first byte in the top nibble.  The grid is sized from BYTES (grid-stride
// loop), not from the number of pairs as in the reference (gasal_align.cu:180-194), so long-sequence
// batches still fill the machine.  HBM-bound: 1 B read + 0.5 B written per base.
__device__ __forceinline__ uint32_t pack8(uint32_t lo4, uint32_t hi4) {
  // lo4 holds bases 0..3 (byte 0 = base 0), hi4 bases 4..7
  uint32_t x = lo4 & 0x0F0F0F0Fu, y = hi4 & 0x0F0F0F0Fu;
  uint32_t xa = ((x & 0x0000000Fu) << 12) | ((x & 0x00000F00u)) | ((x & 0x000F0000u) >> 12) | ((x & 0x0F000000u) >> 24);
  uint32_t ya = ((y & 0x0000000Fu) << 12) | ((y & 0x00000F00u)) | ((y & 0x000F0000u) >> 12) | ((y & 0x0F000000u) >> 24);
  return (xa << 16) | ya;
}

__global__ void __launch_bounds__(256) k_pack(const uint4* __restrict__ in_q, uint2* __restrict__ out_q, uint32_t n16_q,
                                              const uint4* __restrict__ in_t, uint2* __restrict__ out_t, uint32_t n16_t) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t total = n16_q + n16_t;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const bool isq = i < n16_q;
    const uint32_t k = isq ? i : i - n16_q;
    const uint4 v = __ldg((isq ? in_q : in_t) + k);
    uint2 o; o.x = pack8(v.x, v.y); o.y = pack8(v.z, v.w);
    (isq ? out_q : out_t)[k] = o;
  }
}
// tail: batches whose byte count is 8 mod 16 leave one word over
__global__ void k_pack_tail(const uint2* __restrict__ in8, uint32_t* __restrict__ out) {
  const uint2 v = *in8; *out = pack8(v.x, v.y);
}

// ---- K1b: reverse / complement ------------------------------------------------------------------
// op bit0 reverses, op bit1 complements A<->T (1<->4), C<->G (3<->7).  The reference's reverse branch has
// undefined behaviour (32-bit shifts by 32): mode 0 = clean nibble reversal of the pair's whole padded word
// array (PTX shift semantics, the default), mode 1 = what its nvcc 12.9 / sm_100 build actually does.
// One WARP per sequence (lanes stride over the words) instead of one thread per pair.
__device__ __forceinline__ uint32_t complement_word(uint32_t w) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t c = (w >> (4 * k)) & 15u;
    c = (c == 1u) ? 4u : (c == 4u) ? 1u : (c == 3u) ? 7u : (c == 7u) ? 3u : c;
    o |= c << (4 * k);
  }
  return o;
}

__global__ void __launch_bounds__(256) k_revcomp(uint32_t* __restrict__ pq, uint32_t* __restrict__ pt,
                                                 const uint32_t* __restrict__ q_lens, const uint32_t* __restrict__ t_lens,
                                                 const uint32_t* __restrict__ q_offs, const uint32_t* __restrict__ t_offs,
                                                 const uint8_t* __restrict__ q_op, const uint8_t* __restrict__ t_op, uint32_t n, int mode) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * n) return;
  const uint32_t pair = warp >> 1; const bool isq = (warp & 1) == 0;
  const uint8_t op = isq ? q_op[pair] : t_op[pair];
  if (op == 0) return;
  uint32_t* w = (isq ? pq : pt) + ((isq ? q_offs : t_offs)[pair] >> 3);
  const uint32_t nw = n_words(isq ? q_lens[pair] : t_lens[pair]);
  if ((op & 1) && mode == 1) {
    // literal behaviour of the reference as compiled by nvcc 12.9 for sm_100 (undefined shifts folded to 0):
    // out[nw-1-i] = reverse(in[i]), out[i] = 0 for i < ceil(nw/2); complement of a 0 nibble is 0
    for (uint32_t i = lane; i < (nw + 1) / 2; i += 32) {
      uint32_t a = nibrev(w[i]);
      if (op & 2) a = complement_word(a);
      const uint32_t j = nw - 1 - i;
      if (j != i) w[i] = 0u;
      w[j] = a;
    }
    if ((op & 2) && lane == 0) { /* second-half words were all overwritten above; nothing else to complement */ }
  } else if (op & 1) {
    for (uint32_t i = lane; i < (nw + 1) / 2; i += 32) {
      const uint32_t j = nw - 1 - i;
      uint32_t a = nibrev(w[i]), b = nibrev(w[j]);
      if (op & 2) { a = complement_word(a); b = complement_word(b); }
      w[i] = b; w[j] = a;                      // i == j (middle word): both writes store the same value
    }
  } else {
    for (uint32_t i = lane; i < nw; i += 32) w[i] = complement_word(w[i]);
  }
}

// ---- K4: traceback walker -------------------------------------------------------------------------
// Literal state machine of gasal_get_tb (get_tb.h:50-145): bytes (count<<2)|op, op 0=M 1=X 2=D 3=I,
// runs capped at 63, written in REVERSE alignment order at cigar[q_offs[pair]...]; LOCAL stops when
// the re-accumulated score reaches aln_score and reports the start cell; GLOBAL starts one cell
// past the end (Q5) and appends the residual D / I runs.
template <bool IS_LOCAL>
__global__ void __launch_bounds__(128) k_get_tb(const DpArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (uint32_t)a.n) return;
  int i, j, total = 0, curr = 0;
  const uint32_t qlen = a.q_lens[tid], tlen = a.t_lens[tid];
  if (IS_LOCAL) { i = a.t_end[tid]; j = a.q_end[tid]; total = a.score[tid]; }
  else { i = (int)tlen; j = (int)qlen; }
  const int q8 = (int)(n_words(qlen) * 8);
  const int oe = a.gap_open + a.gap_extend;
  const BandGeom bg = band_geom(q8 >> 3, (int)n_words(tlen), a.k_band);   // used when only the tile band is stored
  uint8_t* out = a.cigar + a.q_offs[tid];
  uint32_t prev_op = 0, op_fill = 0;
  int n_ops = 0, count = 0, op_select = 3, op_shift = 0, prev_tile = -1;
  uint4 tile = make_uint4(0, 0, 0, 0);
  while (i >= 0 && j >= 0) {
    int tile_no, in_tile;
    if (IS_LOCAL && a.tb_banded) {
      const int strip = i >> 3, word = j >> 3, lo = band_lo(bg, strip);
      in_tile = ((j & 3) << 3) + (i & 7);
      if (word < lo || word >= band_hi(bg, strip, q8 >> 3)) {
        // outside the stored band (only reachable from a score-0 end cell or through quirk Q14): reads as an all-zero tile
        tile_no = -2; if (prev_tile != -2) { tile = make_uint4(0, 0, 0, 0); prev_tile = -2; }
      } else tile_no = ((strip * bg.W + (word - lo)) << 1) + ((j & 7) >> 2);
    } else {
      const int cell = (((i >> 3) * q8) << 3) + (j << 3) + (i & 7);
      tile_no = cell >> 5;
      in_tile = cell - (tile_no << 5);
    }
    if (tile_no != prev_tile) { tile = a.tb_off ? a.tb[a.tb_off[tid] + (uint64_t)tile_no] : a.tb[(size_t)tile_no * a.tb_stride + tid]; prev_tile = tile_no; }
    const int reg_no = in_tile >> 3, cell_in_reg = in_tile & 7;
    const uint32_t reg = reg_no == 0 ? tile.x : (reg_no == 1 ? tile.y : (reg_no == 2 ? tile.z : tile.w));
    const uint32_t cell_op = (reg >> (28 - (cell_in_reg << 2))) & 15u;
    const uint32_t op = (cell_op >> op_shift) & (uint32_t)op_select;
    op_fill = (op == 0 || op_select == 3) ? op : (uint32_t)op_shift;
    op_select = (op == 0 || (op == 1 && op_select == 3)) ? 3 : 1;
    op_shift = (op == 0 || (op == 1 && op_select == 3)) ? 0 : ((op == 2 || op == 3) ? (int)op : op_shift);
    if (count < 63 && op_fill == prev_op) count++;
    else {
      if (count > 0) { *out++ = (uint8_t)(prev_op | (uint32_t)(count << 2)); n_ops++; }
      count = 1;
    }
    if (IS_LOCAL) {
      curr += ((op_fill == 2 || op_fill == 3) && prev_op != op_fill) ? -oe
            : ((op_fill == 2 || op_fill == 3) ? -a.gap_extend : (op_fill == 1 ? -a.mismatch : a.match));
      if (curr == total) break;
    }
    prev_op = op_fill;
    if (op_fill != 3) i--;
    if (op_fill != 2) j--;
  }
  *out++ = (uint8_t)(prev_op | (uint32_t)(count << 2)); n_ops++;
  if (!IS_LOCAL) {
    while (i >= 0) { const int c = (i + 1) <= 63 ? (i + 1) : 63; *out++ = (uint8_t)(2 | (c << 2)); n_ops++; i -= 63; }
    while (j >= 0) { const int c = (j + 1) <= 63 ? (j + 1) : 63; *out++ = (uint8_t)(3 | (c << 2)); n_ops++; j -= 63; }
  } else { a.t_start[tid] = i; a.q_start[tid] = j; }
  a.n_cigar_ops[tid] = (uint32_t)n_ops;
}

// ---- K5: CIGAR text on the device -----------------------------------------------------------------
// The reference leaves text formatting to a serial host loop (test_prog.cpp:382-428: walk the (count<<2)|op bytes from
// the LAST one to the first, merge neighbouring runs of the same op, print "<count><op>").  Here one thread per pair does
// the same walk on the device and writes the text of pair p at text[p * stride]; lens[p] = characters written.
// style 0: the reference's letters M / X / D / I.   style 1: SAM CIGAR -- M and X merge into M, and for LOCAL the clipped
// query ends become soft clips (q_start leading, q_len - 1 - q_end trailing).  If the text does not fit in stride - 1
// characters it is truncated and lens[p] = 0x80000000 | needed length.
__device__ __forceinline__ void cigar_emit(char* out, uint32_t cap, uint32_t& len, uint32_t count, char op) {
  char d[10]; int nd = 0;
  do { d[nd++] = (char)('0' + count % 10u); count /= 10u; } while (count);
  while (nd > 0) { --nd; if (len < cap) out[len] = d[nd]; len++; }
  if (len < cap) out[len] = op;
  len++;
}

__global__ void __launch_bounds__(128) k_cigar_text(const uint8_t* __restrict__ cigar, const uint32_t* __restrict__ n_cigar_ops,
                                                     const uint32_t* __restrict__ q_offs, const uint32_t* __restrict__ q_lens,
                                                     const int32_t* __restrict__ q_start, const int32_t* __restrict__ q_end,
                                                     int n, int style, int clip, char* __restrict__ text, uint32_t stride, uint32_t* __restrict__ lens) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (uint32_t)n) return;
  const uint8_t* c = cigar + q_offs[tid];
  const uint32_t nops = n_cigar_ops[tid];
  char* out = text + (size_t)tid * stride;
  const uint32_t cap = stride ? stride - 1 : 0;
  uint32_t len = 0;
  const char opc[4] = {'M', style ? 'M' : 'X', 'D', 'I'};
  if (style && clip && q_start[tid] > 0) cigar_emit(out, cap, len, (uint32_t)q_start[tid], 'S');
  if (nops > 0) {
    char last = opc[c[nops - 1] & 3];
    uint32_t count = c[nops - 1] >> 2;
    for (int u = (int)nops - 2; u >= 0; u--) {
      const char op = opc[c[u] & 3];
      if (op == last) count += c[u] >> 2;
      else { cigar_emit(out, cap, len, count, last); count = c[u] >> 2; last = op; }
    }
    cigar_emit(out, cap, len, count, last);
  }
  if (style && clip) { const int tail = (int)q_lens[tid] - 1 - q_end[tid]; if (tail > 0) cigar_emit(out, cap, len, (uint32_t)tail, 'S'); }
  if (stride) out[len < cap ? len : cap] = 0;
  lens[tid] = len <= cap ? len : (0x80000000u | len);
}

}  // namespace gb200
