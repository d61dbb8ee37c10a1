// host_pack.cpp -- 4-bit packing on the HOST, in the device layout of the reference's gasal_pack_kernel
// (Non-CDP/GASAL2/src/kernels/pack_rc_seqs.h:13-53: nibble = byte & 15, first base of a word in bits 31:28).
//
// Why it exists: the reference ships 1 byte per base over PCIe and packs on the GPU (gasal_align.cu:152-201).  On an 8-GPU box
// that is 8 x 1.5 GB per 45 ms step and the host link, not the GPUs, sets the pace (round-1 SCALE: e2e efficiency 0.64).  With
// host packing switched on (the default, gasalx_set_host_packing) gasal_host_batch_fill / _add pack while they copy into the
// pinned page, so the page and the H2D copy carry 0.5 byte per base; the caller still hands over ASCII.
//
// 32 bases per AVX2 iteration: AND 0x0F, PMADDUBSW with (16, 1) merges neighbouring nibbles into bytes, PACKUSWB narrows,
// PSHUFB reverses the four bytes of every word (first base = most significant nibble).  Scalar fallback for other CPUs.
#include <stdint.h>
#include <string.h>
#include <immintrin.h>
#include "host_pack.h"

namespace gb200 {

static inline uint32_t pack8_scalar(const uint8_t* b) {
  return ((uint32_t)(b[0] & 15) << 28) | ((uint32_t)(b[1] & 15) << 24) | ((uint32_t)(b[2] & 15) << 20) | ((uint32_t)(b[3] & 15) << 16) |
         ((uint32_t)(b[4] & 15) << 12) | ((uint32_t)(b[5] & 15) << 8) | ((uint32_t)(b[6] & 15) << 4) | (uint32_t)(b[7] & 15);
}

static void pack_words_scalar(const uint8_t* src, uint64_t nw, uint32_t* dst) {
  for (uint64_t w = 0; w < nw; w++) dst[w] = pack8_scalar(src + 8 * w);
}

__attribute__((target("avx2"))) static void pack_words_avx2(const uint8_t* src, uint64_t nw, uint32_t* dst) {
  const __m256i lo4 = _mm256_set1_epi8(0x0F), mul = _mm256_set1_epi16(0x0110);
  const __m256i rev = _mm256_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12, 3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12);
  uint64_t w = 0;
  for (; w + 4 <= nw; w += 4) {
    const __m256i x = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(src + 8 * w)), lo4);
    const __m256i m = _mm256_maddubs_epi16(x, mul);                 // 16-bit lanes: base[2i] * 16 + base[2i+1]
    const __m256i p = _mm256_shuffle_epi8(_mm256_packus_epi16(m, m), rev);
    const __m256i q = _mm256_permute4x64_epi64(p, 0x08);            // qword 0 of each 128-bit lane -> low 128 bits
    _mm_storeu_si128((__m128i*)(dst + w), _mm256_castsi256_si128(q));
  }
  for (; w < nw; w++) dst[w] = pack8_scalar(src + 8 * w);
}

// ---- "does the data hold a base outside A/C/G/T?" while packing (nibbles 1, 3, 4, 7 are A, C, T, G) --------------------------
static inline bool odd_nibble(uint32_t n) { n &= 15u; return !(n == 1u || n == 3u || n == 4u || n == 7u); }

static bool pack_words_scalar_odd(const uint8_t* src, uint64_t nw, uint32_t* dst) {
  bool odd = false;
  for (uint64_t w = 0; w < nw; w++) {
    const uint8_t* b = src + 8 * w;
    for (int k = 0; k < 8; k++) odd |= odd_nibble(b[k]);
    dst[w] = pack8_scalar(b);
  }
  return odd;
}

__attribute__((target("avx2"))) static bool pack_words_avx2_odd(const uint8_t* src, uint64_t nw, uint32_t* dst) {
  const __m256i lo4 = _mm256_set1_epi8(0x0F), mul = _mm256_set1_epi16(0x0110);
  const __m256i rev = _mm256_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12, 3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12);
  // table over the 16 nibble values: 0 for A/C/G/T (1, 3, 7, 4), 0xFF otherwise
  const __m256i bad = _mm256_setr_epi8(-1, 0, -1, 0, 0, -1, -1, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, -1, 0, 0, -1, -1, 0, -1, -1, -1, -1, -1, -1, -1, -1);
  __m256i acc = _mm256_setzero_si256();
  uint64_t w = 0;
  for (; w + 4 <= nw; w += 4) {
    const __m256i x = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(src + 8 * w)), lo4);
    acc = _mm256_or_si256(acc, _mm256_shuffle_epi8(bad, x));
    const __m256i m = _mm256_maddubs_epi16(x, mul);
    const __m256i p = _mm256_shuffle_epi8(_mm256_packus_epi16(m, m), rev);
    const __m256i q = _mm256_permute4x64_epi64(p, 0x08);
    _mm_storeu_si128((__m128i*)(dst + w), _mm256_castsi256_si128(q));
  }
  bool odd = !_mm256_testz_si256(acc, acc);
  for (; w < nw; w++) { const uint8_t* b = src + 8 * w; for (int k = 0; k < 8; k++) odd |= odd_nibble(b[k]); dst[w] = pack8_scalar(b); }
  return odd;
}

static bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

void pack_words(const uint8_t* src, uint64_t nw, uint32_t* dst) {
  if (have_avx2()) pack_words_avx2(src, nw, dst); else pack_words_scalar(src, nw, dst);
}

static inline void set_nibble(uint32_t* words, uint64_t pos, uint32_t code) {
  const int sh = 28 - 4 * (int)(pos & 7);
  uint32_t& w = words[pos >> 3];
  w = (w & ~(0xFu << sh)) | ((code & 15u) << sh);
}

void pack_bases_at(uint32_t* words, uint64_t pos, const uint8_t* src, uint64_t n) {
  while (n && (pos & 7)) { set_nibble(words, pos, *src); pos++; src++; n--; }
  const uint64_t nw = n >> 3;
  if (nw) { pack_words(src, nw, words + (pos >> 3)); pos += 8 * nw; src += 8 * nw; n -= 8 * nw; }
  while (n) { set_nibble(words, pos, *src); pos++; src++; n--; }
}

bool pack_bases_at_odd(uint32_t* words, uint64_t pos, const uint8_t* src, uint64_t n) {
  bool odd = false;
  while (n && (pos & 7)) { odd |= odd_nibble(*src); set_nibble(words, pos, *src); pos++; src++; n--; }
  const uint64_t nw = n >> 3;
  if (nw) {
    odd |= have_avx2() ? pack_words_avx2_odd(src, nw, words + (pos >> 3)) : pack_words_scalar_odd(src, nw, words + (pos >> 3));
    pos += 8 * nw; src += 8 * nw; n -= 8 * nw;
  }
  while (n) { odd |= odd_nibble(*src); set_nibble(words, pos, *src); pos++; src++; n--; }
  return odd;
}

void fill_nibbles_at(uint32_t* words, uint64_t pos, uint64_t n, uint32_t code) {
  while (n && (pos & 7)) { set_nibble(words, pos, code); pos++; n--; }
  const uint32_t full = (code & 15u) * 0x11111111u;
  for (; n >= 8; n -= 8, pos += 8) words[pos >> 3] = full;
  while (n) { set_nibble(words, pos, code); pos++; n--; }
}

}  // namespace gb200

extern "C" void gasalx_pack_host(const uint8_t* ascii, uint64_t n_bytes, uint32_t* words) { gb200::pack_words(ascii, n_bytes / 8, words); }
extern "C" void gasalx_pack_host_at(uint32_t* words, uint64_t pos, const uint8_t* ascii, uint64_t n) { gb200::pack_bases_at(words, pos, ascii, n); }
