// engine.cu -- host runtime of the B200 alignment engine behind the GASAL2 C API
// (include/gasal_b200.h).  One storage == one CUDA stream; everything gasal_aln_async enqueues
// goes on that stream and nothing synchronises (reference Non-CDP/GASAL2/src/gasal_align.cu:29-307,
// src/ctors.cpp:26-190, src/res.cpp, src/host_batch.cpp, src/interfaces.cpp).
//
// There is NO CPU fallback anywhere in this file: if the CUDA runtime is unusable the calls print
// "[GASAL CUDA ERROR:]" and exit, exactly like the reference's CHECKCUDAERROR (src/gasal.h:15-22).
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../include/gasal_b200.h"
#include "dp_common.cuh"
#include "dp_launch.h"
#include "host_pack.h"

using namespace gb200;

// ---- error handling: the reference prints to stderr and exit()s (src/gasal.h:15-22, gasal_align.cu:32-67); that stays the
// default.  The gasalx_try_* entry points ("opt-in non-fatal status code") run the same code with
// tl_soft set: gasal_die() then throws instead of exiting, the extern "C" wrapper catches and returns the code, and the
// message stays readable through gasalx_last_error().
namespace {
thread_local bool tl_soft = false;
thread_local char tl_msg[512] = "";
struct SoftError { int code; };
[[noreturn]] void gasal_die(int code) {
  if (tl_soft) throw SoftError{code};
  exit(EXIT_FAILURE);
}
}  // namespace
#define GASAL_FAIL(code, ...)                          \
  do {                                                 \
    snprintf(tl_msg, sizeof(tl_msg), __VA_ARGS__);     \
    fputs(tl_msg, stderr);                             \
    gasal_die(code);                                   \
  } while (0)

#define CHECKCUDAERROR(call)                                                                                          \
  do {                                                                                                                \
    cudaError_t err_ = (call);                                                                                        \
    if (cudaSuccess != err_) {                                                                                        \
      GASAL_FAIL(err_ == cudaErrorMemoryAllocation ? GASALX_E_NOMEM : GASALX_E_CUDA,                                    \
                 "[GASAL CUDA ERROR:] %s(CUDA error no.=%d). Line no. %d in file %s\n", cudaGetErrorString(err_),      \
                 (int)err_, __LINE__, __FILE__);                                                                      \
    }                                                                                                                 \
  } while (0)

#define CHECK_LAUNCH() CHECKCUDAERROR(cudaGetLastError())

namespace {

// process-wide scoring, set by gasal_copy_subst_scores (the reference keeps these in __constant__
// memory, src/gasal_kernels.h:29-33; here they travel as kernel arguments)
gasal_subst_scores g_scores = {1, 4, 6, 1};
int g_n_code = 0x4E;
int g_n_penalty = -1;
int g_kernel_policy = 0;
int g_revcomp_mode = 0;
int g_coord_fix = 0;
int g_host_pack = 1;      // gasalx_set_host_packing: gasal_host_batch_fill / _add pack to 4 bits while they copy (0.5 B per base over PCIe)
int g_host_wc = 0;
int g_fast_passes = 0;    // gasalx_set_fast_path_passes: 1 = always run the classification / N-aware / general passes (testing)        // gasalx_set_host_write_combined: sequence pages in write-combined pinned memory

struct Priv {
  uint32_t* colstate = nullptr; size_t colstate_words = 0;
  uint8_t* d_cigar = nullptr; size_t cigar_bytes = 0;
  uint32_t* d_n_cigar_ops = nullptr; size_t n_cigar_cap = 0;
  FastScratch fast;
  WaveScratch wave;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // start, packed, dp done, end
  bool timing_pending = false;
  uint64_t launches = 0;
  const char* last_kernel = "none";
  int max_query_len_8 = 0, max_target_len_8 = 0;
  char* d_text = nullptr; size_t text_bytes = 0;             // gasalx_cigar_text_async
  uint32_t* d_text_lens = nullptr; size_t text_lens_cap = 0;
  uint32_t tb_n = 0; int tb_algo = 0;                         // pairs / algo of the last WITH_TB batch (0: none)
  bool host_packed = false;                                   // the sequence pages hold 4-bit words (latched from g_host_pack at init)
  // host-packed pages: did the packer see a base outside A/C/G/T since the pages were last reset?  A clean batch skips the
  // classification kernel and the N-aware / general passes of the s16x2 fast path.  (`saved`: state before the last reset, for rearm)
  bool host_odd = false, host_odd_saved = false;
  // mixed-class batches (per-pair kernel choice): pairs regrouped by length class; [order | q_lens | t_lens | q_offs | t_offs] staged in
  // pinned memory and mirrored on the device, followed by five compact result arrays
  uint32_t* mix_host = nullptr; uint32_t* mix_dev = nullptr; size_t mix_cap = 0;
  FastScratch fast_mix; WaveScratch wave_mix; uint32_t* colstate_mix = nullptr; size_t colstate_mix_words = 0;
  char kernel_name_buf[48] = "";
  uint32_t fix_n = 0; int fix_kind = 0;                       // gasalx_set_coordinate_fix: 1 local start mirror (Q3), 2 / 3 semi-global default coordinates (Q8; 3: with start)
};

inline Priv* priv(gasal_gpu_storage_t* s) { return (Priv*)s->b200_priv; }

template <typename T>
void dev_realloc(T** p, size_t n) {
  if (*p) CHECKCUDAERROR(cudaFree(*p));
  *p = nullptr;
  if (n) CHECKCUDAERROR(cudaMalloc((void**)p, n * sizeof(T)));
}

inline uint32_t round8(uint32_t x) { return (x + 7u) & ~7u; }

template <typename T>
T* cudaHostRealloc(void* source, int new_size, int old_size) {
  T* destination = NULL;
  if (new_size < old_size) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR] cudoHostRealloc: invalid sizes. New size < old size (%d < %d)", new_size, old_size); }
  CHECKCUDAERROR(cudaHostAlloc(&destination, new_size * sizeof(T), cudaHostAllocDefault));
  memcpy(destination, source, old_size * sizeof(T));
  CHECKCUDAERROR(cudaFreeHost(source));
  return destination;
}


}  // namespace

// =================================================================================================
extern "C" {

const char* gasalx_version(void) { return "gasal-b200 0.1 (sm_100a)"; }

void gasal_params_default(gasal_params_t* p) {   // src/args_parser.cpp:12-29
  memset(p, 0, sizeof(*p));
  p->algo = UNKNOWN; p->start_pos = WITHOUT_START; p->secondBest = FALSE; p->isPacked = 0; p->isReverseComplement = 0;
  p->semiglobal_skipping_head = TARGET; p->semiglobal_skipping_tail = TARGET; p->k_band = 0;
}

void gasalx_set_n_code(int n_code, int n_penalty) { g_n_code = n_code; g_n_penalty = n_penalty; }
void gasalx_set_kernel_policy(int policy) { g_kernel_policy = policy; }
void gasalx_set_revcomp_mode(int mode) { g_revcomp_mode = mode; }
void gasalx_set_coordinate_fix(int on) { g_coord_fix = on; }
void gasalx_set_host_packing(int on) { g_host_pack = on ? 1 : 0; }
void gasalx_set_host_write_combined(int on) { g_host_wc = on ? 1 : 0; }
void gasalx_set_fast_path_passes(int always) { g_fast_passes = always ? 1 : 0; }
int gasalx_storage_host_packed(const gasal_gpu_storage_t* s) { const Priv* pv = (const Priv*)s->b200_priv; return (pv && pv->host_packed) ? 1 : 0; }

void gasal_copy_subst_scores(gasal_subst_scores* subst) { g_scores = *subst; }   // src/gasal_align.cu:329-339

// ---- results: src/res.cpp --------------------------------------------------------------------------
gasal_res_t* gasal_res_new_host(uint32_t max_n_alns, const gasal_params_t* params) {
  gasal_res_t* res = (gasal_res_t*)calloc(1, sizeof(gasal_res_t));   // zero-initialised (the reference mallocs: Q12)
  if (res == NULL) { fprintf(stderr, "Malloc error on res host "); exit(1); }
  CHECKCUDAERROR(cudaHostAlloc(&(res->aln_score), max_n_alns * sizeof(int32_t), cudaHostAllocDefault));
  if (params->algo != GLOBAL) {                                       // :26-30
    CHECKCUDAERROR(cudaHostAlloc(&(res->query_batch_end), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaHostAlloc(&(res->target_batch_end), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
    if (params->start_pos == WITH_START || params->start_pos == WITH_TB) {   // :54-58
      CHECKCUDAERROR(cudaHostAlloc(&(res->query_batch_start), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
      CHECKCUDAERROR(cudaHostAlloc(&(res->target_batch_start), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
    }
  }
  if (params->start_pos == WITH_TB) CHECKCUDAERROR(cudaHostAlloc(&(res->n_cigar_ops), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
  return res;
}

gasal_res_t* gasal_res_new_device_cpy(uint32_t max_n_alns, const gasal_params_t* params) {
  gasal_res_t* res = (gasal_res_t*)calloc(1, sizeof(gasal_res_t));
  CHECKCUDAERROR(cudaMalloc(&(res->aln_score), max_n_alns * sizeof(int32_t)));
  if (params->algo != GLOBAL) {
    CHECKCUDAERROR(cudaMalloc(&(res->query_batch_end), max_n_alns * sizeof(uint32_t)));
    CHECKCUDAERROR(cudaMalloc(&(res->target_batch_end), max_n_alns * sizeof(uint32_t)));
    if (params->start_pos == WITH_START || params->start_pos == WITH_TB) {
      CHECKCUDAERROR(cudaMalloc(&(res->query_batch_start), max_n_alns * sizeof(uint32_t)));
      CHECKCUDAERROR(cudaMalloc(&(res->target_batch_start), max_n_alns * sizeof(uint32_t)));
    }
  }
  return res;
}

gasal_res_t* gasal_res_new_device(gasal_res_t* device_cpy) {   // src/res.cpp:76-101 (one copy instead of five)
  gasal_res_t* d_c;
  CHECKCUDAERROR(cudaMalloc((void**)&d_c, sizeof(gasal_res_t)));
  CHECKCUDAERROR(cudaMemcpy(d_c, device_cpy, sizeof(gasal_res_t), cudaMemcpyHostToDevice));
  return d_c;
}

void gasal_res_destroy_host(gasal_res_t* res) {
  if (res == NULL) return;
  if (res->aln_score != NULL) CHECKCUDAERROR(cudaFreeHost(res->aln_score));
  if (res->query_batch_start != NULL) CHECKCUDAERROR(cudaFreeHost(res->query_batch_start));
  if (res->target_batch_start != NULL) CHECKCUDAERROR(cudaFreeHost(res->target_batch_start));
  if (res->query_batch_end != NULL) CHECKCUDAERROR(cudaFreeHost(res->query_batch_end));
  if (res->target_batch_end != NULL) CHECKCUDAERROR(cudaFreeHost(res->target_batch_end));
  if (res->n_cigar_ops != NULL) CHECKCUDAERROR(cudaFreeHost(res->n_cigar_ops));
  if (res->cigar != NULL) CHECKCUDAERROR(cudaFreeHost(res->cigar));
  free(res);
}

void gasal_res_destroy_device(gasal_res_t* device_res, gasal_res_t* device_cpy) {
  if (device_cpy == NULL || device_res == NULL) return;
  if (device_cpy->aln_score != NULL) CHECKCUDAERROR(cudaFree(device_cpy->aln_score));
  if (device_cpy->query_batch_start != NULL) CHECKCUDAERROR(cudaFree(device_cpy->query_batch_start));
  if (device_cpy->target_batch_start != NULL) CHECKCUDAERROR(cudaFree(device_cpy->target_batch_start));
  if (device_cpy->query_batch_end != NULL) CHECKCUDAERROR(cudaFree(device_cpy->query_batch_end));
  if (device_cpy->target_batch_end != NULL) CHECKCUDAERROR(cudaFree(device_cpy->target_batch_end));
  CHECKCUDAERROR(cudaFree(device_res));
  free(device_cpy);
}

// ---- host staging pages: src/host_batch.cpp ----------------------------------------------------------
// `packed`: the page holds 4-bit words; page_size / data_size / offset stay in BASES (= the reference's bytes), the allocation is half
static host_batch_t* host_page_new(uint32_t batch_bytes, uint32_t offset, bool packed) {
  host_batch_t* res = (host_batch_t*)calloc(1, sizeof(host_batch_t));
  const size_t alloc = packed ? ((size_t)batch_bytes + 15) / 16 * 8 + 8 : (size_t)batch_bytes;
  CHECKCUDAERROR(cudaHostAlloc(&(res->data), alloc, g_host_wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  res->page_size = batch_bytes; res->data_size = 0; res->is_locked = 0; res->offset = offset; res->next = NULL;
  return res;
}

host_batch_t* gasal_host_batch_new(uint32_t batch_bytes, uint32_t offset) { return host_page_new(batch_bytes, offset, false); }

// Host-packed storages keep ONE page per side that grows by reallocation (a linked page could start in the middle of a word).
static void packed_page_reserve(gasal_gpu_storage_t* s, host_batch_t* page, uint32_t* p_batch_bytes, uint64_t need, enum data_source SRC) {
  if (need <= page->page_size) return;
  uint64_t new_size = (uint64_t)page->page_size * 2;
  while (new_size < need) new_size *= 2;
  if (new_size > 0xFFFFFFF0ull) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] host batch of %llu bytes does not fit the 32-bit byte offsets of the GASAL2 API\n", (unsigned long long)need); }
  fprintf(stderr, "[GASAL WARNING:] Trying to write up to byte %llu while the %s host page holds %d bytes. Reallocating it with %llu bytes. Doing this repeadtedly slows down the execution.\n",
          (unsigned long long)need, (SRC == QUERY ? "query" : "target"), page->page_size, (unsigned long long)new_size);
  uint8_t* nd = NULL;
  CHECKCUDAERROR(cudaHostAlloc(&nd, (size_t)(new_size + 15) / 16 * 8 + 8, g_host_wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  memcpy(nd, page->data, ((size_t)page->data_size + 15) / 16 * 8);
  CHECKCUDAERROR(cudaFreeHost(page->data));
  page->data = nd; page->page_size = (uint32_t)new_size;
  *p_batch_bytes = (uint32_t)new_size;
  (void)s;
}

static uint32_t packed_page_write(gasal_gpu_storage_t* s, uint32_t idx, const char* data, uint32_t size, uint32_t pad, enum data_source SRC) {
  host_batch_t* page = (SRC == QUERY) ? s->extensible_host_unpacked_query_batch : s->extensible_host_unpacked_target_batch;
  uint32_t* p_batch_bytes = (SRC == QUERY) ? &s->host_max_query_batch_bytes : &s->host_max_target_batch_bytes;
  const uint64_t end = (uint64_t)idx + size + pad;
  packed_page_reserve(s, page, p_batch_bytes, end, SRC);
  if (gb200::pack_bases_at_odd((uint32_t*)page->data, idx, (const uint8_t*)data, size)) priv(s)->host_odd = true;
  if (pad) gb200::fill_nibbles_at((uint32_t*)page->data, (uint64_t)idx + size, pad, (uint32_t)g_n_code & 15u);
  if (page->data_size < end) page->data_size = (uint32_t)end;
  return (uint32_t)end;
}

void gasal_host_batch_destroy(host_batch_t* res) {
  if (res == NULL) { fprintf(stderr, "[GASAL ERROR] Trying to free a NULL pointer\n"); exit(1); }
  while (res != NULL) {   // iterative (the reference recurses)
    host_batch_t* next = res->next;
    if (res->data != NULL) CHECKCUDAERROR(cudaFreeHost(res->data));
    free(res);
    res = next;
  }
}

host_batch_t* gasal_host_batch_getlast(host_batch_t* arg) { while (arg->next != NULL) arg = arg->next; return arg; }

void gasal_host_batch_reset(gasal_gpu_storage_t* gpu_storage) {   // :49-75
  if (priv(gpu_storage)) { priv(gpu_storage)->host_odd_saved = priv(gpu_storage)->host_odd; priv(gpu_storage)->host_odd = false; }
  host_batch_t* heads[2] = {gpu_storage->extensible_host_unpacked_query_batch, gpu_storage->extensible_host_unpacked_target_batch};
  for (int i = 0; i < 2; i++)
    for (host_batch_t* cur = heads[i]; cur != NULL; cur = cur->next) { cur->data_size = 0; cur->offset = 0; cur->is_locked = 0; }
}

// append one sequence, pad with N_CODE to a multiple of 8, grow by a linked 2x page when full (:79-153)
uint32_t gasal_host_batch_fill(gasal_gpu_storage_t* gpu_storage, uint32_t idx, const char* data, uint32_t size, enum data_source SRC) {
  host_batch_t* cur_page = NULL; uint32_t* p_batch_bytes = NULL;
  switch (SRC) {
    case QUERY: cur_page = gpu_storage->extensible_host_unpacked_query_batch; p_batch_bytes = &(gpu_storage->host_max_query_batch_bytes); break;
    case TARGET: cur_page = gpu_storage->extensible_host_unpacked_target_batch; p_batch_bytes = &(gpu_storage->host_max_target_batch_bytes); break;
    default: GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] gasal_host_batch_fill: SRC must be QUERY or TARGET\n");
  }
  const uint32_t nbr_N = (8u - (size & 7u)) & 7u;
  if (priv(gpu_storage) && priv(gpu_storage)->host_packed) return packed_page_write(gpu_storage, idx, data, size, nbr_N, SRC);
  while (cur_page->is_locked) cur_page = cur_page->next;
  if (cur_page->next == NULL && cur_page->page_size - cur_page->data_size < size + nbr_N) {
    uint32_t new_size = cur_page->page_size * 2;
    while (new_size < size + nbr_N) new_size *= 2;   // (the reference assumes one doubling is enough)
    fprintf(stderr, "[GASAL WARNING:] Trying to write %d bytes while only %d remain (%s) (block size %d, filled %d bytes).\n"
            "                 Allocating a new block of size %d, total size available reaches %d. Doing this repeadtedly slows down the execution.\n",
            size + nbr_N, cur_page->page_size - cur_page->data_size, (SRC == QUERY ? "query" : "target"), cur_page->page_size,
            cur_page->data_size, new_size, *p_batch_bytes + new_size);
    host_batch_t* res = gasal_host_batch_new(new_size, cur_page->offset + cur_page->data_size);
    cur_page->next = res; cur_page->is_locked = 1;
    *p_batch_bytes = *p_batch_bytes + new_size;
    cur_page = cur_page->next;
  }
  if (cur_page->next != NULL && cur_page->page_size - cur_page->data_size < size + nbr_N) {
    cur_page->next->offset = cur_page->offset + cur_page->data_size;
    cur_page->is_locked = 1;
    cur_page = cur_page->next;
  }
  if (cur_page->page_size - cur_page->data_size >= size + nbr_N) {
    memcpy(&(cur_page->data[idx - cur_page->offset]), data, size);
    memset(&(cur_page->data[idx + size - cur_page->offset]), g_n_code, nbr_N);
    idx = idx + size + nbr_N;
    cur_page->data_size += size + nbr_N;
  }
  return idx;
}

uint32_t gasal_host_batch_add(gasal_gpu_storage_t* gpu_storage, uint32_t idx, const char* data, uint32_t size, enum data_source SRC) {   // :163-223
  host_batch_t* cur_page = NULL; uint32_t* p_batch_bytes = NULL;
  switch (SRC) {
    case QUERY: cur_page = gpu_storage->extensible_host_unpacked_query_batch; p_batch_bytes = &(gpu_storage->host_max_query_batch_bytes); break;
    case TARGET: cur_page = gpu_storage->extensible_host_unpacked_target_batch; p_batch_bytes = &(gpu_storage->host_max_target_batch_bytes); break;
    default: GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] gasal_host_batch_add: SRC must be QUERY or TARGET\n");
  }
  if (priv(gpu_storage) && priv(gpu_storage)->host_packed) return packed_page_write(gpu_storage, idx, data, size, 0, SRC);
  int is_done = 0;
  while (!is_done) {
    if (*p_batch_bytes >= idx + size && (cur_page->next == NULL || (cur_page->next->offset >= idx + size))) {
      memcpy(&(cur_page->data[idx - cur_page->offset]), data, size);
      if (cur_page->data_size < idx + size - cur_page->offset) cur_page->data_size = idx + size - cur_page->offset;
      idx = idx + size; is_done = 1;
    } else if ((*p_batch_bytes >= idx + size) && (cur_page->next != NULL) && (cur_page->next->offset < idx + size)) {
      cur_page = cur_page->next;
    } else {
      fprintf(stderr, "[GASAL WARNING:] Trying to write %d bytes at position %d on host memory (%s) while only  %d bytes are available. "
              "Therefore, allocating %d bytes more on CPU. Repeating this many times can provoke a degradation of performance.\n",
              size, idx, (SRC == QUERY ? "query" : "target"), *p_batch_bytes, *p_batch_bytes * 2);
      *p_batch_bytes += *p_batch_bytes;
      while (*p_batch_bytes < size) *p_batch_bytes += *p_batch_bytes;
      host_batch_t* res = gasal_host_batch_new(*p_batch_bytes, idx);
      cur_page->next = res; cur_page = cur_page->next;
    }
  }
  return idx;
}

uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t* gpu_storage, uint32_t idx, const char base, enum data_source SRC) {
  return gasal_host_batch_add(gpu_storage, idx, &base, 1, SRC);
}

void gasal_host_batch_print(host_batch_t* res) {
  fprintf(stderr, "[GASAL PRINT] Page data: offset=%d, next_offset=%d, data size=%d, page size=%d\n", res->offset,
          (res->next != NULL ? (int)res->next->offset : -1), res->data_size, res->page_size);
}
void gasal_host_batch_printall(host_batch_t* res) {
  for (; res != NULL; res = res->next) { gasal_host_batch_print(res); if (res->next) fprintf(stderr, "+--->"); }
}

// ---- src/interfaces.cpp ---------------------------------------------------------------------------------
void gasal_host_alns_resize(gasal_gpu_storage_t* gpu_storage, int new_max_alns, const gasal_params_t* params) {   // :26-78
  fprintf(stderr, "[GASAL WARNING] Resizing gpu_storage from %d sequences to %d sequences... ", gpu_storage->host_max_n_alns, new_max_alns);
  const int old = (int)gpu_storage->host_max_n_alns;
  gpu_storage->host_query_op = cudaHostRealloc<uint8_t>((void*)gpu_storage->host_query_op, new_max_alns, old);
  gpu_storage->host_target_op = cudaHostRealloc<uint8_t>((void*)gpu_storage->host_target_op, new_max_alns, old);
  memset(gpu_storage->host_query_op + old, 0, new_max_alns - old);
  memset(gpu_storage->host_target_op + old, 0, new_max_alns - old);
  if (params->algo == KSW) gpu_storage->host_seed_scores = cudaHostRealloc<uint32_t>(gpu_storage->host_seed_scores, new_max_alns, old);
  gpu_storage->host_query_batch_lens = cudaHostRealloc<uint32_t>((void*)gpu_storage->host_query_batch_lens, new_max_alns, old);
  gpu_storage->host_target_batch_lens = cudaHostRealloc<uint32_t>((void*)gpu_storage->host_target_batch_lens, new_max_alns, old);
  gpu_storage->host_query_batch_offsets = cudaHostRealloc<uint32_t>((void*)gpu_storage->host_query_batch_offsets, new_max_alns, old);
  gpu_storage->host_target_batch_offsets = cudaHostRealloc<uint32_t>((void*)gpu_storage->host_target_batch_offsets, new_max_alns, old);
  uint8_t* old_cigar = gpu_storage->host_res->cigar; gpu_storage->host_res->cigar = NULL;   // the CIGAR page is sized in bytes, keep it
  gasal_res_destroy_host(gpu_storage->host_res);
  gpu_storage->host_res = gasal_res_new_host(new_max_alns, params);
  gpu_storage->host_res->cigar = old_cigar;
  if (params->secondBest) {
    gasal_res_destroy_host(gpu_storage->host_res_second);
    gpu_storage->host_res_second = gasal_res_new_host(new_max_alns, params);
  }
  // device side is regrown by gasal_aln_async when the batch is larger than gpu_max_n_alns
  gpu_storage->host_max_n_alns = new_max_alns;
  fprintf(stderr, " done. This can harm performance.\n");
}

void gasal_op_fill(gasal_gpu_storage_t* gpu_storage_t, uint8_t* data, uint32_t nbr_seqs_in_stream, enum data_source SRC) {   // :81-96
  uint8_t* host_op = NULL;
  switch (SRC) { case QUERY: host_op = gpu_storage_t->host_query_op; break; case TARGET: host_op = gpu_storage_t->host_target_op; break; default: break; }
  if (host_op) memcpy(host_op, data, nbr_seqs_in_stream);
}

void gasal_set_device(int gpu_select, int isPrintingProp) {   // :98-128
  if (isPrintingProp) {
    int num_devices = 0;
    cudaGetDeviceCount(&num_devices);
    fprintf(stderr, "Found %d GPUs\n", num_devices);
    if (gpu_select > num_devices - 1) {
      GASAL_FAIL(GASALX_E_ARG, "Error: can't select device %d when only %d devices are selected (range from 0 to %d)\n", gpu_select, num_devices, num_devices - 1);
    }
    for (int device = 0; device < num_devices; device++) {
      cudaDeviceProp properties; cudaGetDeviceProperties(&properties, device);
      fprintf(stderr, "\tGPU %d: %s\n", device, properties.name);
    }
    if (num_devices > 0) {
      cudaDeviceProp properties; cudaGetDeviceProperties(&properties, gpu_select);
      fprintf(stderr, "Selected device %d : %s\n", gpu_select, properties.name);
    }
  }
  CHECKCUDAERROR(cudaSetDevice(gpu_select));
}

// ---- src/ctors.cpp -----------------------------------------------------------------------------------------
gasal_gpu_storage_v gasal_init_gpu_storage_v(int n_streams) {
  gasal_gpu_storage_v v;
  v.a = (gasal_gpu_storage_t*)calloc(n_streams, sizeof(gasal_gpu_storage_t));
  v.n = n_streams;
  return v;
}

static void alloc_tb(gasal_gpu_storage_t* s, uint64_t q8, uint64_t t8, uint64_t n) {
  // ctors.cpp:112-115, plus one strip of slack so the GLOBAL walker's first read (Q5) stays in bounds
  const uint64_t tiles = (q8 * (t8 + 8) + 31) / 32 + 2;
  if (tiles * n > s->packed_tb_matrix_size) {
    if (s->packed_tb_matrices) CHECKCUDAERROR(cudaFree(s->packed_tb_matrices));
    s->packed_tb_matrix_size = tiles * n;
    CHECKCUDAERROR(cudaMalloc(&(s->packed_tb_matrices), s->packed_tb_matrix_size * sizeof(uint4)));
  }
}

void gasal_init_streams(gasal_gpu_storage_v* gpu_storage_vec, int max_query_len, int max_target_len, int max_n_alns, const gasal_params_t* params) {
  const int max_query_len_8 = (int)round8((uint32_t)max_query_len), max_target_len_8 = (int)round8((uint32_t)max_target_len);
  const uint64_t qb = (uint64_t)max_n_alns * max_query_len_8, tb = (uint64_t)max_n_alns * max_target_len_8;
  if (qb > 0xFFFFFFF0ull || tb > 0xFFFFFFF0ull) {
    GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] batch of %d alignments x (%d, %d) bases does not fit the 32-bit byte offsets of the GASAL2 API\n", max_n_alns, max_query_len_8, max_target_len_8);
  }
  for (int i = 0; i < gpu_storage_vec->n; i++) {
    gasal_gpu_storage_t* s = &gpu_storage_vec->a[i];
    Priv* pv = new Priv();
    s->b200_priv = pv;
    pv->max_query_len_8 = max_query_len_8; pv->max_target_len_8 = max_target_len_8;
    // host packing (default): the pinned pages and the H2D copies carry 4-bit words, the device needs no ASCII buffers at all
    pv->host_packed = g_host_pack && !params->isPacked;
    s->extensible_host_unpacked_query_batch = host_page_new((uint32_t)qb, 0, pv->host_packed);
    s->extensible_host_unpacked_target_batch = host_page_new((uint32_t)tb, 0, pv->host_packed);
    CHECKCUDAERROR(cudaMalloc(&(s->unpacked_query_batch), pv->host_packed ? qb / 2 : qb));
    CHECKCUDAERROR(cudaMalloc(&(s->unpacked_target_batch), pv->host_packed ? tb / 2 : tb));
    CHECKCUDAERROR(cudaHostAlloc(&(s->host_query_op), max_n_alns, cudaHostAllocDefault));
    CHECKCUDAERROR(cudaHostAlloc(&(s->host_target_op), max_n_alns, cudaHostAllocDefault));
    memset(s->host_query_op, 0, max_n_alns); memset(s->host_target_op, 0, max_n_alns);
    CHECKCUDAERROR(cudaMalloc(&(s->query_op), max_n_alns));
    CHECKCUDAERROR(cudaMalloc(&(s->target_op), max_n_alns));
    if (params->isPacked || pv->host_packed) {   // ctors.cpp:64-67: the "unpacked" buffers ARE the packed words
      s->packed_query_batch = (uint32_t*)s->unpacked_query_batch;
      s->packed_target_batch = (uint32_t*)s->unpacked_target_batch;
    } else {
      CHECKCUDAERROR(cudaMalloc(&(s->packed_query_batch), (qb / 8) * sizeof(uint32_t)));
      CHECKCUDAERROR(cudaMalloc(&(s->packed_target_batch), (tb / 8) * sizeof(uint32_t)));
    }
    if (params->algo == KSW) {
      CHECKCUDAERROR(cudaHostAlloc(&(s->host_seed_scores), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
      CHECKCUDAERROR(cudaMalloc(&(s->seed_scores), max_n_alns * sizeof(uint32_t)));
    }
    CHECKCUDAERROR(cudaHostAlloc(&(s->host_query_batch_lens), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaHostAlloc(&(s->host_target_batch_lens), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaHostAlloc(&(s->host_query_batch_offsets), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaHostAlloc(&(s->host_target_batch_offsets), max_n_alns * sizeof(uint32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaMalloc(&(s->query_batch_lens), max_n_alns * sizeof(uint32_t)));
    CHECKCUDAERROR(cudaMalloc(&(s->target_batch_lens), max_n_alns * sizeof(uint32_t)));
    CHECKCUDAERROR(cudaMalloc(&(s->query_batch_offsets), max_n_alns * sizeof(uint32_t)));
    CHECKCUDAERROR(cudaMalloc(&(s->target_batch_offsets), max_n_alns * sizeof(uint32_t)));
    s->host_res = gasal_res_new_host(max_n_alns, params);
    if (params->start_pos == WITH_TB) CHECKCUDAERROR(cudaHostAlloc(&(s->host_res->cigar), qb, cudaHostAllocDefault));
    s->device_cpy = gasal_res_new_device_cpy(max_n_alns, params);
    s->device_res = gasal_res_new_device(s->device_cpy);
    if (params->secondBest) {
      s->host_res_second = gasal_res_new_host(max_n_alns, params);
      s->device_cpy_second = gasal_res_new_device_cpy(max_n_alns, params);
      s->device_res_second = gasal_res_new_device(s->device_cpy_second);
    }
    if (params->start_pos == WITH_TB && (params->algo == LOCAL || params->algo == GLOBAL))
      alloc_tb(s, max_query_len_8, max_target_len_8, max_n_alns);   // (BANDED + WITH_TB sizes its band matrix per batch)
    CHECKCUDAERROR(cudaStreamCreateWithFlags(&(s->str), cudaStreamNonBlocking));
    for (int e = 0; e < 4; e++) CHECKCUDAERROR(cudaEventCreate(&pv->ev[e]));
    s->is_free = 1;
    s->host_max_query_batch_bytes = (uint32_t)qb; s->host_max_target_batch_bytes = (uint32_t)tb;
    s->host_max_n_alns = max_n_alns;
    s->gpu_max_query_batch_bytes = (uint32_t)qb; s->gpu_max_target_batch_bytes = (uint32_t)tb;
    s->gpu_max_n_alns = max_n_alns;
    s->current_n_alns = 0;
  }
}

void gasal_destroy_streams(gasal_gpu_storage_v* gpu_storage_vec, const gasal_params_t* params) {
  for (int i = 0; i < gpu_storage_vec->n; i++) {
    gasal_gpu_storage_t* s = &gpu_storage_vec->a[i];
    if (s->str) CHECKCUDAERROR(cudaStreamSynchronize(s->str));
    gasal_host_batch_destroy(s->extensible_host_unpacked_query_batch);
    gasal_host_batch_destroy(s->extensible_host_unpacked_target_batch);
    gasal_res_destroy_host(s->host_res);
    gasal_res_destroy_device(s->device_res, s->device_cpy);
    gasal_res_destroy_host(s->host_res_second);
    gasal_res_destroy_device(s->device_res_second, s->device_cpy_second);
    if (s->seed_scores) CHECKCUDAERROR(cudaFree(s->seed_scores));
    if (s->host_seed_scores) CHECKCUDAERROR(cudaFreeHost(s->host_seed_scores));
    if (s->query_op) CHECKCUDAERROR(cudaFree(s->query_op));
    if (s->target_op) CHECKCUDAERROR(cudaFree(s->target_op));
    if (s->host_query_op) CHECKCUDAERROR(cudaFreeHost(s->host_query_op));
    if (s->host_target_op) CHECKCUDAERROR(cudaFreeHost(s->host_target_op));
    if (s->host_query_batch_offsets) CHECKCUDAERROR(cudaFreeHost(s->host_query_batch_offsets));
    if (s->host_target_batch_offsets) CHECKCUDAERROR(cudaFreeHost(s->host_target_batch_offsets));
    if (s->host_query_batch_lens) CHECKCUDAERROR(cudaFreeHost(s->host_query_batch_lens));
    if (s->host_target_batch_lens) CHECKCUDAERROR(cudaFreeHost(s->host_target_batch_lens));
    if (s->unpacked_query_batch) CHECKCUDAERROR(cudaFree(s->unpacked_query_batch));
    if (s->unpacked_target_batch) CHECKCUDAERROR(cudaFree(s->unpacked_target_batch));
    if ((void*)s->packed_query_batch != (void*)s->unpacked_query_batch && s->packed_query_batch) CHECKCUDAERROR(cudaFree(s->packed_query_batch));
    if ((void*)s->packed_target_batch != (void*)s->unpacked_target_batch && s->packed_target_batch) CHECKCUDAERROR(cudaFree(s->packed_target_batch));
    if (s->query_batch_offsets) CHECKCUDAERROR(cudaFree(s->query_batch_offsets));
    if (s->target_batch_offsets) CHECKCUDAERROR(cudaFree(s->target_batch_offsets));
    if (s->query_batch_lens) CHECKCUDAERROR(cudaFree(s->query_batch_lens));
    if (s->target_batch_lens) CHECKCUDAERROR(cudaFree(s->target_batch_lens));
    if (s->packed_tb_matrices) CHECKCUDAERROR(cudaFree(s->packed_tb_matrices));
    Priv* pv = priv(s);
    if (pv) {
      if (pv->colstate) CHECKCUDAERROR(cudaFree(pv->colstate));
      if (pv->d_cigar) CHECKCUDAERROR(cudaFree(pv->d_cigar));
      if (pv->d_n_cigar_ops) CHECKCUDAERROR(cudaFree(pv->d_n_cigar_ops));
      if (pv->d_text) CHECKCUDAERROR(cudaFree(pv->d_text));
      if (pv->d_text_lens) CHECKCUDAERROR(cudaFree(pv->d_text_lens));
      pv->fast.release();
      pv->wave.release();
      pv->fast_mix.release(); pv->wave_mix.release();
      if (pv->colstate_mix) CHECKCUDAERROR(cudaFree(pv->colstate_mix));
      if (pv->mix_host) CHECKCUDAERROR(cudaFreeHost(pv->mix_host));
      if (pv->mix_dev) CHECKCUDAERROR(cudaFree(pv->mix_dev));
      for (int e = 0; e < 4; e++) if (pv->ev[e]) CHECKCUDAERROR(cudaEventDestroy(pv->ev[e]));
      delete pv;
    }
    if (s->str) CHECKCUDAERROR(cudaStreamDestroy(s->str));
    memset(s, 0, sizeof(*s));
  }
}

// Deprecated in the reference (ctors.cpp:202-252): device buffers of ONE storage, no host pages, no stream.
void gasal_gpu_mem_alloc(gasal_gpu_storage_t* s, int gpu_max_query_batch_bytes, int gpu_max_target_batch_bytes, int gpu_max_n_alns, const gasal_params_t* params) {
  (void)params;
  CHECKCUDAERROR(cudaMalloc(&(s->unpacked_query_batch), (size_t)gpu_max_query_batch_bytes));
  CHECKCUDAERROR(cudaMalloc(&(s->unpacked_target_batch), (size_t)gpu_max_target_batch_bytes));
  CHECKCUDAERROR(cudaMalloc(&(s->packed_query_batch), ((size_t)gpu_max_query_batch_bytes / 8) * sizeof(uint32_t)));
  CHECKCUDAERROR(cudaMalloc(&(s->packed_target_batch), ((size_t)gpu_max_target_batch_bytes / 8) * sizeof(uint32_t)));
  CHECKCUDAERROR(cudaMalloc(&(s->query_batch_lens), (size_t)gpu_max_n_alns * sizeof(uint32_t)));
  CHECKCUDAERROR(cudaMalloc(&(s->target_batch_lens), (size_t)gpu_max_n_alns * sizeof(uint32_t)));
  CHECKCUDAERROR(cudaMalloc(&(s->query_batch_offsets), (size_t)gpu_max_n_alns * sizeof(uint32_t)));
  CHECKCUDAERROR(cudaMalloc(&(s->target_batch_offsets), (size_t)gpu_max_n_alns * sizeof(uint32_t)));
  if (s->device_cpy) s->device_res = gasal_res_new_device(s->device_cpy);
  s->gpu_max_query_batch_bytes = (uint32_t)gpu_max_query_batch_bytes; s->gpu_max_target_batch_bytes = (uint32_t)gpu_max_target_batch_bytes;
  s->gpu_max_n_alns = (uint32_t)gpu_max_n_alns;
}

void gasal_gpu_mem_free(gasal_gpu_storage_t* s, const gasal_params_t* params) {
  const bool alias_q = (void*)s->packed_query_batch == (void*)s->unpacked_query_batch, alias_t = (void*)s->packed_target_batch == (void*)s->unpacked_target_batch;
  if (s->unpacked_query_batch != NULL) CHECKCUDAERROR(cudaFree(s->unpacked_query_batch));
  if (s->unpacked_target_batch != NULL) CHECKCUDAERROR(cudaFree(s->unpacked_target_batch));
  if (!alias_q && s->packed_query_batch != NULL) CHECKCUDAERROR(cudaFree(s->packed_query_batch));
  if (!alias_t && s->packed_target_batch != NULL) CHECKCUDAERROR(cudaFree(s->packed_target_batch));
  if (s->query_batch_offsets != NULL) CHECKCUDAERROR(cudaFree(s->query_batch_offsets));
  if (s->target_batch_offsets != NULL) CHECKCUDAERROR(cudaFree(s->target_batch_offsets));
  if (s->query_batch_lens != NULL) CHECKCUDAERROR(cudaFree(s->query_batch_lens));
  if (s->target_batch_lens != NULL) CHECKCUDAERROR(cudaFree(s->target_batch_lens));
  s->unpacked_query_batch = s->unpacked_target_batch = NULL; s->packed_query_batch = s->packed_target_batch = NULL;
  s->query_batch_offsets = s->target_batch_offsets = s->query_batch_lens = s->target_batch_lens = NULL;
  gasal_res_destroy_device(s->device_res, s->device_cpy); s->device_res = s->device_cpy = NULL;
  if (params->secondBest) { gasal_res_destroy_device(s->device_res_second, s->device_cpy_second); s->device_res_second = s->device_cpy_second = NULL; }
}

void gasal_destroy_gpu_storage_v(gasal_gpu_storage_v* gpu_storage_vec) {
  if (gpu_storage_vec->a != NULL) free(gpu_storage_vec->a);
  gpu_storage_vec->a = NULL; gpu_storage_vec->n = 0;
}

// ---- the hot path -----------------------------------------------------------------------------------------------
static void grow_device(gasal_gpu_storage_t* s, uint32_t q_bytes, uint32_t t_bytes, uint32_t n, const gasal_params_t* params) {
  // src/gasal_align.cu:70-145 (same policy: integer multiples, warn on stderr)
  if (s->gpu_max_query_batch_bytes < q_bytes) {
    int i = 2; while ((uint64_t)s->gpu_max_query_batch_bytes * i < q_bytes) i++;
    fprintf(stderr, "[GASAL WARNING:] actual_query_batch_bytes(%d) > Allocated GPU memory (gpu_max_query_batch_bytes=%d). Therefore, allocating %d bytes on GPU (gpu_max_query_batch_bytes=%d). Performance may be lost if this is repeated many times.\n",
            q_bytes, s->gpu_max_query_batch_bytes, s->gpu_max_query_batch_bytes * i, s->gpu_max_query_batch_bytes * i);
    s->gpu_max_query_batch_bytes *= i;
    const bool alias = (void*)s->packed_query_batch == (void*)s->unpacked_query_batch;
    dev_realloc(&s->unpacked_query_batch, priv(s)->host_packed ? s->gpu_max_query_batch_bytes / 2 : s->gpu_max_query_batch_bytes);
    if (alias) s->packed_query_batch = (uint32_t*)s->unpacked_query_batch;
    else dev_realloc(&s->packed_query_batch, s->gpu_max_query_batch_bytes / 8);
    if (params->start_pos == WITH_TB) {
      if (s->host_res->cigar != NULL) CHECKCUDAERROR(cudaFreeHost(s->host_res->cigar));
      CHECKCUDAERROR(cudaHostAlloc(&(s->host_res->cigar), s->gpu_max_query_batch_bytes, cudaHostAllocDefault));
    }
  }
  if (s->gpu_max_target_batch_bytes < t_bytes) {
    int i = 2; while ((uint64_t)s->gpu_max_target_batch_bytes * i < t_bytes) i++;
    fprintf(stderr, "[GASAL WARNING:] actual_target_batch_bytes(%d) > Allocated GPU memory (gpu_max_target_batch_bytes=%d). Therefore, allocating %d bytes on GPU (gpu_max_target_batch_bytes=%d). Performance may be lost if this is repeated many times.\n",
            t_bytes, s->gpu_max_target_batch_bytes, s->gpu_max_target_batch_bytes * i, s->gpu_max_target_batch_bytes * i);
    s->gpu_max_target_batch_bytes *= i;
    const bool alias = (void*)s->packed_target_batch == (void*)s->unpacked_target_batch;
    dev_realloc(&s->unpacked_target_batch, priv(s)->host_packed ? s->gpu_max_target_batch_bytes / 2 : s->gpu_max_target_batch_bytes);
    if (alias) s->packed_target_batch = (uint32_t*)s->unpacked_target_batch;
    else dev_realloc(&s->packed_target_batch, s->gpu_max_target_batch_bytes / 8);
  }
  if (s->gpu_max_n_alns < n) {
    int i = 2; while ((uint64_t)s->gpu_max_n_alns * i < n) i++;
    fprintf(stderr, "[GASAL WARNING:] actual_n_alns(%d) > gpu_max_n_alns(%d). Therefore, allocating memory for %d alignments on  GPU (gpu_max_n_alns=%d). Performance may be lost if this is repeated many times.\n",
            n, s->gpu_max_n_alns, s->gpu_max_n_alns * i, s->gpu_max_n_alns * i);
    s->gpu_max_n_alns *= i;
    dev_realloc(&s->query_batch_lens, s->gpu_max_n_alns); dev_realloc(&s->target_batch_lens, s->gpu_max_n_alns);
    dev_realloc(&s->query_batch_offsets, s->gpu_max_n_alns); dev_realloc(&s->target_batch_offsets, s->gpu_max_n_alns);
    dev_realloc(&s->query_op, s->gpu_max_n_alns); dev_realloc(&s->target_op, s->gpu_max_n_alns);
    if (s->seed_scores) dev_realloc(&s->seed_scores, s->gpu_max_n_alns);
    gasal_res_destroy_device(s->device_res, s->device_cpy);
    s->device_cpy = gasal_res_new_device_cpy(s->gpu_max_n_alns, params);
    s->device_res = gasal_res_new_device(s->device_cpy);
    if (params->secondBest) {
      gasal_res_destroy_device(s->device_res_second, s->device_cpy_second);
      s->device_cpy_second = gasal_res_new_device_cpy(s->gpu_max_n_alns, params);
      s->device_res_second = gasal_res_new_device(s->device_cpy_second);
    }
  }
}

static int device_sms() { int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); return sms; }

static void launch_pack(gasal_gpu_storage_t* s, uint32_t q_bytes, uint32_t t_bytes) {
  Priv* pv = priv(s);
  pv->launches += gb200::launch_pack(s->unpacked_query_batch, s->packed_query_batch, q_bytes, s->unpacked_target_batch, s->packed_target_batch, t_bytes, device_sms(), s->str);
  CHECK_LAUNCH();
}

struct BatchShape { uint32_t max_q, max_t, min_q, min_t; };

// ---- per-pair kernel choice (one submission, several kernel families) ---------------------------------------------------------
// The reference runs every pair of a batch on the same one-thread-per-pair kernel.  Here the kernel family depends on the pair's
// shape, and a plain gasal_aln_async caller may mix shapes freely: when a batch holds both short queries (s16x2 inter-task fast
// path) and long ones (intra-task wavefront kernel), the pairs are regrouped by class on the host (the lengths are in the
// caller's pinned arrays anyway), each class runs on its own kernel family over compact copies of the lens / offsets arrays, and
// one small kernel scatters the compact results back to the caller's pair order.  The sequences themselves do not move.
// Returns false when the batch is not mixed (or the mode has no mixed path): the caller then dispatches the whole batch as before.
static bool launch_dp_mixed(gasal_gpu_storage_t* s, uint32_t n, const gasal_params_t* params, const DpArgs& a0, const BatchShape& shape) {
  Priv* pv = priv(s);
  const int algo = params->algo, start = params->start_pos;
  if (g_kernel_policy != 0 || params->secondBest || start == WITH_TB) return false;
  if (!(algo == LOCAL || algo == SEMI_GLOBAL || algo == GLOBAL || algo == BANDED)) return false;      // (BANDED: short class on the s16x2 banded sweep, long class on the int32 banded kernel)
  if (round8(shape.max_q) <= kFastMaxQ8 || round8(shape.min_q) > kFastMaxQ8 || n < 2) return false;   // all short or all long
  const uint32_t* ql = s->host_query_batch_lens; const uint32_t* tl = s->host_target_batch_lens;
  BatchShape sh[2];
  uint32_t cnt[2] = {0, 0};
  for (int c = 0; c < 2; c++) { sh[c].max_q = sh[c].max_t = 0; sh[c].min_q = sh[c].min_t = 0xFFFFFFFFu; }
  for (uint32_t i = 0; i < n; i++) {
    const int c = round8(ql[i]) <= kFastMaxQ8 ? 0 : 1;
    cnt[c]++;
    sh[c].max_q = std::max(sh[c].max_q, ql[i]); sh[c].max_t = std::max(sh[c].max_t, tl[i]);
    sh[c].min_q = std::min(sh[c].min_q, ql[i]); sh[c].min_t = std::min(sh[c].min_t, tl[i]);
  }
  if (cnt[0] == 0 || cnt[1] == 0) return false;
  if (!fast_path_eligible(a0, algo, start, false, sh[0].max_q, sh[0].max_t)) return false;
  // staging: [order | q_lens | t_lens | q_offs | t_offs] (5n words, pinned + device), then 5n words of compact results on the device
  if (pv->mix_cap < n) {
    if (pv->mix_host) CHECKCUDAERROR(cudaFreeHost(pv->mix_host));
    if (pv->mix_dev) CHECKCUDAERROR(cudaFree(pv->mix_dev));
    pv->mix_host = nullptr; pv->mix_dev = nullptr;
    CHECKCUDAERROR(cudaHostAlloc(&pv->mix_host, 5 * (size_t)n * sizeof(uint32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaMalloc(&pv->mix_dev, 10 * (size_t)n * sizeof(uint32_t)));
    pv->mix_cap = n;
  }
  const size_t N = pv->mix_cap;
  uint32_t* h = pv->mix_host;
  uint32_t at[2] = {0, cnt[0]};
  for (uint32_t i = 0; i < n; i++) {
    const int c = round8(ql[i]) <= kFastMaxQ8 ? 0 : 1;
    const uint32_t k = at[c]++;
    h[k] = i; h[N + k] = ql[i]; h[2 * N + k] = tl[i]; h[3 * N + k] = s->host_query_batch_offsets[i]; h[4 * N + k] = s->host_target_batch_offsets[i];
  }
  uint32_t* d = pv->mix_dev;
  for (int k = 0; k < 5; k++) CHECKCUDAERROR(cudaMemcpyAsync(d + k * N, h + k * N, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, s->str));
  int32_t* res[5];
  for (int k = 0; k < 5; k++) res[k] = (int32_t*)(d + (5 + k) * N);
  int32_t* dst[5] = {a0.score, a0.q_end, a0.t_end, a0.q_start, a0.t_start};
  const char* names[2] = {"s16x2", "int32"};
  for (int c = 0; c < 2; c++) {
    const uint32_t off = c ? cnt[0] : 0;
    DpArgs a = a0;
    a.n = (int)cnt[c];
    a.q_lens = d + N + off; a.t_lens = d + 2 * N + off; a.q_offs = d + 3 * N + off; a.t_offs = d + 4 * N + off;
    a.score = res[0] + off; a.q_end = dst[1] ? res[1] + off : nullptr; a.t_end = dst[2] ? res[2] + off : nullptr;
    a.q_start = dst[3] ? res[3] + off : nullptr; a.t_start = dst[4] ? res[4] + off : nullptr;
    if (c == 0) {
      const int nl = fast_path_launch(a, algo, start, pv->fast_mix, sh[0].max_q, sh[0].max_t, !(pv->host_packed && !pv->host_odd), g_fast_passes, s->str);
      if (nl < 0) { GASAL_FAIL(GASALX_E_NOMEM, "[GASAL CUDA ERROR:] out of device memory for the fast-path scratch\n"); }
      CHECK_LAUNCH(); pv->launches += nl;
    } else if (wave_eligible(a, algo, start, false, sh[1].min_q, sh[1].min_t)) {
      bool s16 = wave16_eligible(a, algo, sh[1].max_q, sh[1].max_t);
      const int nl = wave_launch(a, algo, start, pv->wave_mix, h + N + off, h + 2 * N + off, sh[1].max_t, device_sms(), s->str, &s16, false);
      if (nl < 0) { GASAL_FAIL(GASALX_E_NOMEM, "[GASAL CUDA ERROR:] out of device memory for the wavefront scratch\n"); }
      CHECK_LAUNCH(); pv->launches += nl; names[1] = s16 ? "wave16" : "wave";
    } else {
      const uint32_t q8 = round8(sh[1].max_q), stride = (cnt[1] + 31u) & ~31u;
      const size_t need = (size_t)stride * std::max<uint32_t>(q8, 8);
      if (pv->colstate_mix_words < need) { dev_realloc(&pv->colstate_mix, need); pv->colstate_mix_words = need; }
      a.colstate = pv->colstate_mix; a.col_stride = stride; a.col_rows = q8;
      const int nl = generic_launch(a, algo, start, false, s->str);
      if (nl < 0) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] algo %d is not supported by this build\n", algo); }
      CHECK_LAUNCH(); pv->launches += nl;
    }
  }
  for (int k = 0; k < 5; k++) if (!dst[k]) res[k] = nullptr;
  launch_scatter_results(d, n, res, dst, s->str); CHECK_LAUNCH(); pv->launches++;
  snprintf(pv->kernel_name_buf, sizeof(pv->kernel_name_buf), "mixed(%s+%s)", names[0], names[1]);
  pv->last_kernel = pv->kernel_name_buf;
  return true;
}

static void launch_dp(gasal_gpu_storage_t* s, uint32_t q_bytes, uint32_t n, const gasal_params_t* params, const BatchShape& shape) {
  Priv* pv = priv(s);
  DpArgs a; memset(&a, 0, sizeof(a));
  a.pq = s->packed_query_batch; a.pt = s->packed_target_batch;
  a.q_lens = s->query_batch_lens; a.t_lens = s->target_batch_lens; a.q_offs = s->query_batch_offsets; a.t_offs = s->target_batch_offsets;
  a.seed_scores = s->seed_scores; a.n = (int)n;
  a.match = g_scores.match; a.mismatch = g_scores.mismatch; a.gap_open = g_scores.gap_open; a.gap_extend = g_scores.gap_extend;
  a.n_value = (uint32_t)g_n_code & 15u; a.n_score = (g_n_penalty >= 0) ? -g_n_penalty : 0; a.n_is_wild_global = (g_n_penalty >= 0);
  a.score = s->device_cpy->aln_score; a.q_end = s->device_cpy->query_batch_end; a.t_end = s->device_cpy->target_batch_end;
  a.q_start = s->device_cpy->query_batch_start; a.t_start = s->device_cpy->target_batch_start;
  if (s->device_cpy_second) { a.score2 = s->device_cpy_second->aln_score; a.q_end2 = s->device_cpy_second->query_batch_end; a.t_end2 = s->device_cpy_second->target_batch_end; }
  a.head = params->semiglobal_skipping_head; a.tail = params->semiglobal_skipping_tail; a.k_band = params->k_band;
  const uint32_t q8 = round8(shape.max_q), t8 = round8(shape.max_t);
  const bool second = params->secondBest != 0;
  const int start = params->start_pos;
  const bool want_tb = (start == WITH_TB) && (params->algo == LOCAL || params->algo == GLOBAL || params->algo == BANDED);

  // LOCAL WITH_TB on long pairs: the s16x2 wavefront kernel emits the direction matrix into per-pair regions (DpArgs::tb_off)
  const bool wave_tb = want_tb && params->algo == LOCAL && !second && g_kernel_policy != 1 && g_kernel_policy != 3 &&
                       (g_kernel_policy >= 2 || q8 > kFastMaxQ8) && wave_eligible(a, params->algo, start, second, shape.min_q, shape.min_t) &&
                       wave16_eligible(a, params->algo, shape.max_q, shape.max_t);
  if (want_tb) {
    if (wave_tb) {
      uint64_t tiles = 2;
      for (uint32_t i = 0; i < n; i++) tiles += wave_tb_tiles_of(s->host_query_batch_lens[i], s->host_target_batch_lens[i]);
      if (tiles > s->packed_tb_matrix_size) {
        if (s->packed_tb_matrices) CHECKCUDAERROR(cudaFree(s->packed_tb_matrices));
        s->packed_tb_matrices = NULL; s->packed_tb_matrix_size = 0;
        CHECKCUDAERROR(cudaMalloc(&(s->packed_tb_matrices), tiles * sizeof(uint4)));
        s->packed_tb_matrix_size = tiles;
      }
    } else if (params->algo == BANDED) {
      // band matrix: TR strips x W words x 2 tiles, W = widest strip of any pair in the batch (dp_common.cuh band_geom)
      const int qr_max = (int)(q8 >> 3), qr_min = (int)((shape.min_q + 7) >> 3), tr_max = (int)(t8 >> 3), kw = params->k_band >> 3;
      const int w_max = std::max(1, std::min(qr_max, kw + (tr_max - (qr_min - kw)) - 1));
      const uint64_t tiles = (uint64_t)tr_max * (uint64_t)w_max * 2 + 2;
      if (tiles * n > s->packed_tb_matrix_size) {
        if (s->packed_tb_matrices) CHECKCUDAERROR(cudaFree(s->packed_tb_matrices));
        s->packed_tb_matrix_size = tiles * n;
        CHECKCUDAERROR(cudaMalloc(&(s->packed_tb_matrices), s->packed_tb_matrix_size * sizeof(uint4)));
      }
      a.tb_banded = 1;
    } else alloc_tb(s, q8, t8, n);
    a.tb = s->packed_tb_matrices; a.tb_stride = n;
    if (pv->cigar_bytes < q_bytes) { dev_realloc(&pv->d_cigar, q_bytes); pv->cigar_bytes = q_bytes; }
    if (pv->n_cigar_cap < n) { dev_realloc(&pv->d_n_cigar_ops, n); pv->n_cigar_cap = n; }
    a.cigar = pv->d_cigar; a.n_cigar_ops = pv->d_n_cigar_ops;
  }
  pv->tb_n = want_tb ? n : 0; pv->tb_algo = want_tb ? params->algo : 0;
  pv->fix_kind = 0; pv->fix_n = n;
  if (g_coord_fix) {
    if (params->algo == LOCAL && start == WITH_START) pv->fix_kind = 1;
    else if (params->algo == SEMI_GLOBAL && a.head == TARGET && a.tail == TARGET) pv->fix_kind = (start == WITH_START) ? 3 : 2;
  }

  // ---- batches that mix short and long queries: one kernel family per class ----
  if (launch_dp_mixed(s, n, params, a, shape)) return;

  // ---- fast path: s16x2 DPX kernels (dp_s16.cuh) ----
  if (g_kernel_policy == 0 && fast_path_eligible(a, params->algo, start, second, shape.max_q, shape.max_t)) {
    const int nl = fast_path_launch(a, params->algo, start, pv->fast, shape.max_q, shape.max_t, !(pv->host_packed && !pv->host_odd), g_fast_passes, s->str);
    if (nl < 0) { GASAL_FAIL(GASALX_E_NOMEM, "[GASAL CUDA ERROR:] out of device memory for the fast-path scratch\n"); }
    CHECK_LAUNCH(); pv->launches += nl; pv->last_kernel = "s16x2";
    if (want_tb) { launch_get_tb(a, true, s->str); CHECK_LAUNCH(); pv->launches++; }      // gasal_align.h:45-53
    return;
  }

  if (wave_tb) {
    bool s16 = true;
    const int nl = wave_launch(a, params->algo, start, pv->wave, s->host_query_batch_lens, s->host_target_batch_lens, shape.max_t, device_sms(), s->str, &s16, true);
    if (nl < 0) { GASAL_FAIL(GASALX_E_NOMEM, "[GASAL CUDA ERROR:] out of device memory for the wavefront scratch\n"); }
    CHECK_LAUNCH(); pv->launches += nl; pv->last_kernel = "wave16";
    a.tb_off = pv->wave.toff;                                                             // per-pair regions: the walker addresses them the same way
    launch_get_tb(a, true, s->str); CHECK_LAUNCH(); pv->launches++;                       // gasal_align.h:45-53
    return;
  }

  // ---- long pairs: intra-task wavefront kernel (dp_wave.cuh) ----
  // policy 0: whenever the batch is too long for the fast path; policy 2: wherever the wavefront kernel is eligible
  if (g_kernel_policy != 1 && start != WITH_TB && wave_eligible(a, params->algo, start, second, shape.min_q, shape.min_t) &&
      (g_kernel_policy >= 2 || q8 > kFastMaxQ8)) {
    const int sms = device_sms();
    bool s16 = g_kernel_policy != 3 && wave16_eligible(a, params->algo, shape.max_q, shape.max_t);
    const int nl = wave_launch(a, params->algo, start, pv->wave, s->host_query_batch_lens, s->host_target_batch_lens, shape.max_t, sms, s->str, &s16, g_kernel_policy == 4);
    if (nl < 0) { GASAL_FAIL(GASALX_E_NOMEM, "[GASAL CUDA ERROR:] out of device memory for the wavefront scratch\n"); }
    CHECK_LAUNCH(); pv->launches += nl; pv->last_kernel = s16 ? "wave16" : "wave";
    return;
  }

  // ---- general int32 kernels ----
  const uint32_t stride = (n + 31u) & ~31u;
  pv->last_kernel = "int32";
  int galgo = params->algo;
  if (params->algo == KSW && g_kernel_policy == 0 && s->host_seed_scores) {
    // KSW keeps every stored H / E >= 0, so a u16 row in shared memory is exact while h0 + qlen * match stays below 2^16 (k_ksw_smem):
    // the row of a 150 bp read is 608 B per thread, two CTAs per SM, and none of the 16 B per cell goes through L2 / HBM
    uint32_t max_h0 = 0;
    for (uint32_t i = 0; i < n; i++) max_h0 = std::max(max_h0, s->host_seed_scores[i]);
    const uint64_t top = (uint64_t)max_h0 + (uint64_t)q8 * (uint64_t)std::max(a.match, 0) + 16;
    if (top <= 65535u && a.match >= 0 && a.mismatch >= 0 && a.gap_open >= 0 && a.gap_extend >= 0 && (size_t)(q8 + 2) * 128 * 4 <= (size_t)227 * 1024) {
      galgo = ALGO_KSW_SMEM; pv->last_kernel = "int32-smem";
    }
  }
  // column state in the coalesced global scratch (not needed when the KSW row lives in shared memory)
  const size_t need = (galgo == ALGO_KSW_SMEM) ? 0 : (params->algo == KSW) ? 2 * (size_t)stride * (q8 + 16) : (size_t)stride * std::max<uint32_t>(q8, 8);
  if (pv->colstate_words < need) { dev_realloc(&pv->colstate, need); pv->colstate_words = need; }
  a.colstate = pv->colstate; a.col_stride = stride; a.col_rows = q8;
  const int nl = generic_launch(a, galgo, start, second, s->str);
  if (nl < 0) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] algo %d is not supported by this build\n", params->algo); }
  CHECK_LAUNCH(); pv->launches += nl;
}

static BatchShape scan_shape(const gasal_gpu_storage_t* s, uint32_t n, const Priv* pv, bool have_host_lens) {
  BatchShape sh;
  if (!have_host_lens) { sh.max_q = sh.min_q = pv->max_query_len_8; sh.max_t = sh.min_t = pv->max_target_len_8; return sh; }
  uint32_t mq = 0, mt = 0, nq = 0xFFFFFFFFu, nt = 0xFFFFFFFFu;
  const uint32_t* ql = s->host_query_batch_lens; const uint32_t* tl = s->host_target_batch_lens;
  for (uint32_t i = 0; i < n; i++) { mq = std::max(mq, ql[i]); mt = std::max(mt, tl[i]); nq = std::min(nq, ql[i]); nt = std::min(nt, tl[i]); }
  sh.max_q = mq; sh.max_t = mt; sh.min_q = nq; sh.min_t = nt;
  return sh;
}

static void validate(gasal_gpu_storage_t* s, uint32_t q_bytes, uint32_t t_bytes, uint32_t n) {
  // src/gasal_align.cu:32-67
  if (n <= 0) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_n_alns <= 0\n"); }
  if (q_bytes <= 0) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_query_batch_bytes <= 0\n"); }
  if (t_bytes <= 0) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_target_batch_bytes <= 0\n"); }
  if (q_bytes % 8) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_query_batch_bytes=%d is not a multiple of 8\n", q_bytes); }
  if (t_bytes % 8) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_target_batch_bytes=%d is not a multiple of 8\n", t_bytes); }
  if (q_bytes > s->host_max_query_batch_bytes) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_query_batch_bytes(%d) > host_max_query_batch_bytes(%d)\n", q_bytes, s->host_max_query_batch_bytes); }
  if (t_bytes > s->host_max_target_batch_bytes) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_target_batch_bytes(%d) > host_max_target_batch_bytes(%d)\n", t_bytes, s->host_max_target_batch_bytes); }
  if (n > s->host_max_n_alns) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] actual_n_alns(%d) > host_max_n_alns(%d)\n", n, s->host_max_n_alns); }
}

static void device_stage(gasal_gpu_storage_t* s, uint32_t q_bytes, uint32_t t_bytes, uint32_t n, const gasal_params_t* params, const BatchShape& shape) {
  Priv* pv = priv(s);
  if (!params->isPacked && !pv->host_packed) launch_pack(s, q_bytes, t_bytes);          // gasal_align.cu:190-201 (host-packed pages arrive as words)
  if (params->isReverseComplement) {                                                    // :232-246
    launch_revcomp(s->packed_query_batch, s->packed_target_batch, s->query_batch_lens, s->target_batch_lens,
                   s->query_batch_offsets, s->target_batch_offsets, s->query_op, s->target_op, n, g_revcomp_mode, s->str);
    CHECK_LAUNCH(); pv->launches++;
  }
  CHECKCUDAERROR(cudaEventRecord(pv->ev[1], s->str));
  launch_dp(s, q_bytes, n, params, shape);                                              // :249
  CHECKCUDAERROR(cudaEventRecord(pv->ev[2], s->str));
}

void gasal_aln_async(gasal_gpu_storage_t* s, const uint32_t actual_query_batch_bytes, const uint32_t actual_target_batch_bytes,
                     const uint32_t actual_n_alns, const gasal_params_t* params) {
  validate(s, actual_query_batch_bytes, actual_target_batch_bytes, actual_n_alns);
  Priv* pv = priv(s);
  if (pv == NULL) { GASAL_FAIL(GASALX_E_ARG, "[GASAL ERROR:] storage was not initialised with gasal_init_streams\n"); }
  grow_device(s, actual_query_batch_bytes, actual_target_batch_bytes, actual_n_alns, params);
  const uint32_t n = actual_n_alns;
  const BatchShape shape = scan_shape(s, n, pv, true);

  CHECKCUDAERROR(cudaEventRecord(pv->ev[0], s->str));
  // H2D: one copy per host page (gasal_align.cu:152-175), then lens/offsets (:207-210), seeds, ops
  if (pv->host_packed) {   // one page per side holding 4-bit words: half the bytes, and no pack kernel behind the copy
    CHECKCUDAERROR(cudaMemcpyAsync(s->packed_query_batch, s->extensible_host_unpacked_query_batch->data, actual_query_batch_bytes / 2, cudaMemcpyHostToDevice, s->str));
    CHECKCUDAERROR(cudaMemcpyAsync(s->packed_target_batch, s->extensible_host_unpacked_target_batch->data, actual_target_batch_bytes / 2, cudaMemcpyHostToDevice, s->str));
  } else {
    for (host_batch_t* cur = s->extensible_host_unpacked_query_batch; cur != NULL; cur = cur->next)
      if (cur->data_size) CHECKCUDAERROR(cudaMemcpyAsync(&(s->unpacked_query_batch[cur->offset]), cur->data, cur->data_size, cudaMemcpyHostToDevice, s->str));
    for (host_batch_t* cur = s->extensible_host_unpacked_target_batch; cur != NULL; cur = cur->next)
      if (cur->data_size) CHECKCUDAERROR(cudaMemcpyAsync(&(s->unpacked_target_batch[cur->offset]), cur->data, cur->data_size, cudaMemcpyHostToDevice, s->str));
  }
  CHECKCUDAERROR(cudaMemcpyAsync(s->query_batch_lens, s->host_query_batch_lens, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s->str));
  CHECKCUDAERROR(cudaMemcpyAsync(s->target_batch_lens, s->host_target_batch_lens, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s->str));
  CHECKCUDAERROR(cudaMemcpyAsync(s->query_batch_offsets, s->host_query_batch_offsets, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s->str));
  CHECKCUDAERROR(cudaMemcpyAsync(s->target_batch_offsets, s->host_target_batch_offsets, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s->str));
  if (params->algo == KSW) {
    if (s->seed_scores == NULL || s->host_seed_scores == NULL) { GASAL_FAIL(GASALX_E_ARG, "seed_scores == NULL\n"); }
    CHECKCUDAERROR(cudaMemcpyAsync(s->seed_scores, s->host_seed_scores, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s->str));
  }
  if (params->isReverseComplement) {
    // one-time note: the reference's reverse branch is undefined C (pack_rc_seqs.h:115,136,153); mode 0 is the sane reversal, NOT what
    // its sm_100 binary computes (mode 1 is)
    static bool noted = false;
    if (!noted && g_revcomp_mode == 0) {
      for (uint32_t i = 0; i < n; i++)
        if ((s->host_query_op[i] | s->host_target_op[i]) & 1) {
          fprintf(stderr, "[GASAL NOTE:] a reverse operation is applied with gasalx_set_revcomp_mode(0): sequences are reversed as a whole (padding included); the "
                          "reference library built by nvcc 12.9 for sm_100 zeroes half of every reversed sequence instead -- gasalx_set_revcomp_mode(1) reproduces that bit for bit\n");
          noted = true; break;
        }
    }
    CHECKCUDAERROR(cudaMemcpyAsync(s->query_op, s->host_query_op, n, cudaMemcpyHostToDevice, s->str));
    CHECKCUDAERROR(cudaMemcpyAsync(s->target_op, s->host_target_op, n, cudaMemcpyHostToDevice, s->str));
  }

  device_stage(s, actual_query_batch_bytes, actual_target_batch_bytes, n, params, shape);

  // D2H (gasal_align.cu:266-304)
  gasal_res_t* h = s->host_res; gasal_res_t* d = s->device_cpy;
  if (h->aln_score && d->aln_score) CHECKCUDAERROR(cudaMemcpyAsync(h->aln_score, d->aln_score, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
  if (h->query_batch_start && d->query_batch_start) CHECKCUDAERROR(cudaMemcpyAsync(h->query_batch_start, d->query_batch_start, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
  if (h->target_batch_start && d->target_batch_start) CHECKCUDAERROR(cudaMemcpyAsync(h->target_batch_start, d->target_batch_start, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
  if (h->query_batch_end && d->query_batch_end) CHECKCUDAERROR(cudaMemcpyAsync(h->query_batch_end, d->query_batch_end, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
  if (h->target_batch_end && d->target_batch_end) CHECKCUDAERROR(cudaMemcpyAsync(h->target_batch_end, d->target_batch_end, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
  if (params->start_pos == WITH_TB && pv->d_cigar && h->cigar) {
    CHECKCUDAERROR(cudaMemcpyAsync(h->cigar, pv->d_cigar, actual_query_batch_bytes, cudaMemcpyDeviceToHost, s->str));
    CHECKCUDAERROR(cudaMemcpyAsync(h->n_cigar_ops, pv->d_n_cigar_ops, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->str));
  }
  if (params->secondBest && s->host_res_second && s->device_cpy_second) {
    gasal_res_t* h2 = s->host_res_second; gasal_res_t* d2 = s->device_cpy_second;
    if (h2->aln_score && d2->aln_score) CHECKCUDAERROR(cudaMemcpyAsync(h2->aln_score, d2->aln_score, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
    if (h2->query_batch_end && d2->query_batch_end) CHECKCUDAERROR(cudaMemcpyAsync(h2->query_batch_end, d2->query_batch_end, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
    if (h2->target_batch_end && d2->target_batch_end) CHECKCUDAERROR(cudaMemcpyAsync(h2->target_batch_end, d2->target_batch_end, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->str));
  }
  CHECKCUDAERROR(cudaEventRecord(pv->ev[3], s->str));
  pv->timing_pending = true;
  s->is_free = 0;
}

void gasalx_aln_device_resident(gasal_gpu_storage_t* s, uint32_t q_bytes, uint32_t t_bytes, uint32_t n, const gasal_params_t* params) {
  Priv* pv = priv(s);
  pv->host_odd = pv->host_odd_saved;                              // the device still holds the batch that was uploaded last
  BatchShape shape = scan_shape(s, n, pv, true);
  CHECKCUDAERROR(cudaEventRecord(pv->ev[0], s->str));
  device_stage(s, q_bytes, t_bytes, n, params, shape);
  CHECKCUDAERROR(cudaEventRecord(pv->ev[3], s->str));
  pv->timing_pending = true;
  s->is_free = 0;
}

int gasal_is_aln_async_done(gasal_gpu_storage_t* s) {   // src/gasal_align.cu:310-326
  if (s->is_free == 1) return -2;
  cudaError_t err = cudaStreamQuery(s->str);
  if (err != cudaSuccess) {
    if (err == cudaErrorNotReady) return -1;
    GASAL_FAIL(GASALX_E_CUDA, "[GASAL CUDA ERROR:] %s(CUDA error no.=%d). Line no. %d in file %s\n", cudaGetErrorString(err), (int)err, __LINE__, __FILE__);
  }
  // opt-in repair of two reference quirks in the HOST results (never active in parity runs)
  Priv* pv = priv(s);
  if (pv && pv->fix_kind && s->host_res) {
    gasal_res_t* h = s->host_res;
    if (pv->fix_kind == 1 && h->target_batch_start) {
      // Q3: reported = 16*floor(ts/8) + 14 - ts (the start is mirrored inside its 8-base word, local_kernel_template.h:41)
      for (uint32_t i = 0; i < pv->fix_n; i++) {
        const int r = h->target_batch_start[i] - 7;
        if (r >= 0) h->target_batch_start[i] = (r & ~7) + 7 - (r & 7);
      }
    } else if (pv->fix_kind >= 2 && h->query_batch_end) {
      // Q8: the default semi-global mode aligns the WHOLE query; the reference reports ref_len / read_len-1-ref_len there
      for (uint32_t i = 0; i < pv->fix_n; i++) {
        h->query_batch_end[i] = (int32_t)s->host_query_batch_lens[i] - 1;
        if (pv->fix_kind == 3 && h->query_batch_start) h->query_batch_start[i] = 0;
      }
    }
    pv->fix_kind = 0;
  }
  gasal_host_batch_reset(s);
  s->is_free = 1;
  s->current_n_alns = 0;
  return 0;
}

int gasalx_last_batch_ms(gasal_gpu_storage_t* s, float out_ms[3]) {
  Priv* pv = priv(s);
  if (!pv || !pv->timing_pending) return -1;
  if (cudaEventQuery(pv->ev[3]) != cudaSuccess) return -1;
  CHECKCUDAERROR(cudaEventElapsedTime(&out_ms[0], pv->ev[0], pv->ev[3]));
  CHECKCUDAERROR(cudaEventElapsedTime(&out_ms[2], pv->ev[1], pv->ev[2]));
  float to_packed; CHECKCUDAERROR(cudaEventElapsedTime(&to_packed, pv->ev[0], pv->ev[1]));
  out_ms[1] = to_packed;
  return 0;
}

// SURVEY 8(d) primary time: first H2D enqueue .. last D2H complete over a group of storages of one device that were in flight together
int gasalx_span_ms(gasal_gpu_storage_t* const* storages, int n, float* out_ms) {
  if (n <= 0 || !storages || !out_ms) return -1;
  Priv* p0 = priv(storages[0]);
  if (!p0 || !p0->timing_pending) return -1;
  float lo = 0.f, hi = 0.f;
  for (int i = 0; i < n; i++) {
    Priv* pv = priv(storages[i]);
    if (!pv || !pv->timing_pending || cudaEventQuery(pv->ev[3]) != cudaSuccess) return -1;
    float a = 0.f, b = 0.f;
    CHECKCUDAERROR(cudaEventElapsedTime(&a, p0->ev[0], pv->ev[0]));     // negative when storage i started before storage 0
    CHECKCUDAERROR(cudaEventElapsedTime(&b, p0->ev[0], pv->ev[3]));
    if (i == 0 || a < lo) lo = a;
    if (i == 0 || b > hi) hi = b;
  }
  *out_ms = hi - lo;
  return 0;
}

void gasalx_host_batch_rearm(gasal_gpu_storage_t* s, uint32_t q_bytes, uint32_t t_bytes, uint32_t n_alns) {
  if (priv(s)) priv(s)->host_odd = priv(s)->host_odd_saved;       // the same bytes are still in the pages
  host_batch_t* heads[2] = {s->extensible_host_unpacked_query_batch, s->extensible_host_unpacked_target_batch};
  uint32_t want[2] = {q_bytes, t_bytes};
  for (int i = 0; i < 2; i++) {
    uint32_t off = 0, left = want[i];
    for (host_batch_t* cur = heads[i]; cur != NULL; cur = cur->next) {
      const uint32_t take = left < cur->page_size ? left : cur->page_size;
      cur->offset = off; cur->data_size = take; cur->is_locked = (cur->next != NULL && left > take);
      off += take; left -= take;
    }
  }
  s->current_n_alns = n_alns;
}

int gasalx_cigar_text_async(gasal_gpu_storage_t* s, int style, char* host_text, uint32_t stride, uint32_t* host_lens) {
  Priv* pv = priv(s);
  if (!pv || pv->tb_n == 0 || !pv->d_cigar || !host_text || !host_lens || stride < 2) return -1;
  const uint32_t n = pv->tb_n;
  const size_t bytes = (size_t)n * stride;
  if (pv->text_bytes < bytes) { dev_realloc(&pv->d_text, bytes); pv->text_bytes = bytes; }
  if (pv->text_lens_cap < n) { dev_realloc(&pv->d_text_lens, n); pv->text_lens_cap = n; }
  const bool local_like = (pv->tb_algo == LOCAL || pv->tb_algo == BANDED);        // start/end cells exist: soft clips
  launch_cigar_text(pv->d_cigar, pv->d_n_cigar_ops, s->query_batch_offsets, s->query_batch_lens, s->device_cpy->query_batch_start,
                    s->device_cpy->query_batch_end, (int)n, style, local_like ? 1 : 0, pv->d_text, stride, pv->d_text_lens, s->str);
  CHECK_LAUNCH(); pv->launches++;
  CHECKCUDAERROR(cudaMemcpyAsync(host_text, pv->d_text, bytes, cudaMemcpyDeviceToHost, s->str));
  CHECKCUDAERROR(cudaMemcpyAsync(host_lens, pv->d_text_lens, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->str));
  return 0;
}

int gasalx_try_aln_async(gasal_gpu_storage_t* s, uint32_t actual_query_batch_bytes, uint32_t actual_target_batch_bytes, uint32_t actual_n_alns,
                         const gasal_params_t* params) {
  tl_soft = true; tl_msg[0] = 0;
  int rc = GASALX_OK;
  try { gasal_aln_async(s, actual_query_batch_bytes, actual_target_batch_bytes, actual_n_alns, params); }
  catch (const SoftError& e) { rc = e.code; }
  tl_soft = false;
  return rc;
}

int gasalx_try_init_streams(gasal_gpu_storage_v* gpu_storage_vec, int max_query_len, int max_target_len, int max_n_alns, const gasal_params_t* params) {
  tl_soft = true; tl_msg[0] = 0;
  int rc = GASALX_OK;
  try { gasal_init_streams(gpu_storage_vec, max_query_len, max_target_len, max_n_alns, params); }
  catch (const SoftError& e) { rc = e.code; }
  tl_soft = false;
  return rc;
}

const char* gasalx_last_error(void) { return tl_msg; }

uint64_t gasalx_kernel_launches(const gasal_gpu_storage_t* s) { const Priv* pv = (const Priv*)s->b200_priv; return pv ? pv->launches : 0; }
const char* gasalx_last_kernel_name(const gasal_gpu_storage_t* s) { const Priv* pv = (const Priv*)s->b200_priv; return pv ? pv->last_kernel : "none"; }

}  // extern "C"
