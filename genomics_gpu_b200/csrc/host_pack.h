// host_pack.h -- host-side 4-bit packing helpers (host_pack.cpp); positions and counts are in BASES.
#pragma once
#include <stdint.h>
namespace gb200 {
void pack_words(const uint8_t* src, uint64_t n_words, uint32_t* dst);                      // 8 ASCII bases -> one word, n_words times
void pack_bases_at(uint32_t* words, uint64_t pos, const uint8_t* src, uint64_t n);         // n ASCII bases into nibbles pos .. pos + n - 1
bool pack_bases_at_odd(uint32_t* words, uint64_t pos, const uint8_t* src, uint64_t n);     // same; true if a base outside A/C/G/T was seen
void fill_nibbles_at(uint32_t* words, uint64_t pos, uint64_t n, uint32_t code);            // n copies of one 4-bit code
}
