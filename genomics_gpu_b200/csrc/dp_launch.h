// dp_launch.h -- host-side interface between the runtime (engine.cu) and the kernel translation units.
// Each kernel family is compiled in its own .cu (dp_launch_s16.cu, dp_launch_wave.cu, dp_launch_generic.cu) so that the
// library builds in parallel; engine.cu only sees these declarations.  Reference dispatcher: Non-CDP/GASAL2/src/gasal_align.h:7-107.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "dp_common.cuh"

namespace gb200 {

// ---- s16x2 fast path (dp_s16.cuh) -----------------------------------------------------------------------------------
constexpr uint32_t kFastMaxQ8 = 352;      // (q8 + kSlotPad) * 128 * 5 B must fit one CTA's shared memory

struct FastScratch {
  uint32_t* worklist = nullptr; uint32_t* work_count = nullptr; size_t cap = 0;       // worklist[cap] (pairs with N), worklist2[cap] (general kernel)
  uint32_t* colstate = nullptr; size_t colstate_words = 0;
  uint32_t* flags = nullptr;                                                           // k_classify: pass bits of every pair
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;       // the N-aware launch runs beside the main launch
  bool attr_set = false;
  void release() {
    if (worklist) cudaFree(worklist);
    if (work_count) cudaFree(work_count);
    if (colstate) cudaFree(colstate);
    if (flags) cudaFree(flags);
    if (side) { cudaStreamSynchronize(side); cudaStreamDestroy(side); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    worklist = work_count = colstate = nullptr; flags = nullptr; side = nullptr; ev_fork = ev_join = nullptr; cap = 0; colstate_words = 0;
  }
};

bool fast_path_eligible(const DpArgs& a, int algo, int start, bool second, uint32_t max_q, uint32_t max_t);
// returns the number of kernel launches, or -1 when a scratch allocation failed
// maybe_odd: the batch may hold bases outside A/C/G/T (false: the host packer saw none, the extra passes are skipped)
int fast_path_launch(const DpArgs& a, int algo, int start, FastScratch& fs, uint32_t max_q, uint32_t max_t, bool maybe_odd, int force_passes, cudaStream_t str);

// ---- nvbio-style batched scoring on the quad-strip layout (nvb_strip.cuh; called by nvb_batched.cu) -----------------------
struct NvbStripArgs {
  const uint32_t* poffs; const uint32_t* ppk; const uint8_t* pflag;       // pattern byte offsets, 2-bit packed codes (k_nvb_pack), "not pure ACGT" flags
  const uint32_t* toffs; const uint32_t* tpk; const uint8_t* tflag;
  const uint32_t* pair_p; const uint32_t* pair_t; int text_shared;
  uint32_t first, count;                                                  // job range of this launch
  int match, mismatch, oe, ext;                                           // GASAL2-style magnitudes: mismatch >= 0, oe = -Go, ext = -Ge
  int32_t* score; int32_t* sink_i; int32_t* sink_j;
  uint32_t* worklist; uint32_t* work_count;                               // jobs with a symbol outside A/C/G/T (int32 kernel)
  uint32_t m8max;                                                         // columns reserved per thread in shared memory (multiple of 8)
  uint32_t eight, one;                                                    // == 8, 1 at run time (see FastArgs)
};
// type: 0 GLOBAL, 1 LOCAL, 2 SEMI_GLOBAL (aln::AlignmentType order)
bool nvb_strip_eligible(int type, int match, int mismatch_pen, int oe, int ext, uint32_t max_m, uint32_t max_n, int want_sinks);
int nvb_strip_launch(const NvbStripArgs& a, int type, int want_sinks, cudaStream_t str);

// ---- intra-task wavefront kernels (dp_wave.cuh) ---------------------------------------------------------------------
struct WaveScratch {
  uint64_t* qoff = nullptr; uint64_t* soff = nullptr; uint64_t* toff = nullptr; size_t n_cap = 0;    // device: qoff[n], soff[n], toff[n] (traceback tile offsets)
  uint64_t* h_off = nullptr; size_t h_cap = 0;                    // pinned staging: qoff[n], soff[n], toff[n]
  uint4* bnd = nullptr; size_t bnd_cap = 0;
  uint32_t next_tag = 1;                                          // boundary tags handed out so far (never reused while bnd lives)
  int2* cand = nullptr; size_t cand_cap = 0;
  uint32_t* prog = nullptr; size_t prog_cap = 0;                  // done[n], then the ticket counter
  void release() {
    if (qoff) cudaFree(qoff);
    if (h_off) cudaFreeHost(h_off);
    if (bnd) cudaFree(bnd);
    if (cand) cudaFree(cand);
    if (prog) cudaFree(prog);
    qoff = soff = toff = nullptr; h_off = nullptr; bnd = nullptr; cand = nullptr; prog = nullptr; n_cap = h_cap = bnd_cap = cand_cap = prog_cap = 0;
  }
};

// tiles (uint4) of the per-pair traceback regions of the wavefront LOCAL WITH_TB kernel: pair p owns 2 QR x TR tiles
inline uint64_t wave_tb_tiles_of(uint32_t qlen, uint32_t tlen) { return 2ull * ((qlen + 7u) >> 3) * ((tlen + 7u) >> 3); }
bool wave16_eligible(const DpArgs& a, int algo, uint32_t max_q, uint32_t max_t);
bool wave_eligible(const DpArgs& a, int algo, int start, bool second, uint32_t min_q, uint32_t min_t);
int wave_launch(const DpArgs& a, int algo, int start, WaveScratch& ws, const uint32_t* h_qlens, const uint32_t* h_tlens,
                uint32_t max_t, int n_sm, cudaStream_t str, bool* s16_inout = nullptr, bool force_s16 = false);

// ---- general int32 kernels, pack / reverse-complement, traceback walker, CIGAR text (dp_generic.cuh, pack_tb.cuh) -----
// number of launches, or -1 if the algorithm is not supported
int generic_launch(const DpArgs& a, int algo, int start, bool second, cudaStream_t str);
int launch_pack(const uint8_t* uq, uint32_t* pq, uint32_t q_bytes, const uint8_t* ut, uint32_t* pt, uint32_t t_bytes, int n_sm, cudaStream_t str);
void launch_revcomp(uint32_t* pq, uint32_t* pt, const uint32_t* q_lens, const uint32_t* t_lens, const uint32_t* q_offs, const uint32_t* t_offs,
                    const uint8_t* q_op, const uint8_t* t_op, uint32_t n, int mode, cudaStream_t str);
void launch_get_tb(const DpArgs& a, bool is_local, cudaStream_t str);
// mixed-class batches: dst[k][order[i]] = src[k][i] for the (up to five) result arrays that exist
void launch_scatter_results(const uint32_t* order, uint32_t n, int32_t* const src[5], int32_t* const dst[5], cudaStream_t str);
void launch_cigar_text(const uint8_t* cigar, const uint32_t* n_cigar_ops, const uint32_t* q_offs, const uint32_t* q_lens, const int32_t* q_start,
                       const int32_t* q_end, int n, int style, int clip, char* text, uint32_t stride, uint32_t* lens, cudaStream_t str);

}  // namespace gb200
