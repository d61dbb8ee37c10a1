/*
 * gasal_b200.h -- the C ABI of the B200-native batched pairwise-alignment engine.
 *
 * This is the drop-in boundary for the GASAL2 path of Genomics-GPU (reference tree
 * Non-CDP/GASAL2/src/).  Every type and entry point below keeps the NAME, argument meaning,
 * ownership rules and error behaviour of the reference interface it replaces (cited file:line);
 * the only change is that the reference's C++ `Parameters*` argument (src/args_parser.h:24-68, a
 * class with std::string / std::ifstream members) becomes the POD `gasal_params_t`, so the library
 * is callable from C, cgo, JNI, ctypes ...  C++ callers that want the unchanged reference
 * signatures include gasal_header.h instead, which layers `class Parameters` on top of this file.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 */
#ifndef GASAL_B200_H
#define GASAL_B200_H

#include <stdint.h>
#include <stdlib.h>

#if defined(__CUDACC__) || defined(__CUDA_RUNTIME_H__) || defined(GASAL_B200_WITH_CUDA_RUNTIME)
#include <cuda_runtime.h>
#else
/* C callers without the CUDA headers: same representation as the toolkit's types */
typedef struct CUstream_st* cudaStream_t;
#ifndef GASAL_B200_UINT4_DEFINED
#define GASAL_B200_UINT4_DEFINED
typedef struct __attribute__((aligned(16))) { unsigned int x, y, z, w; } uint4;
#endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums: src/gasal.h:37-72 (same enumerator names and values) ---------------------------- */
enum comp_start { WITHOUT_START, WITH_START, WITH_TB };
enum Bool { FALSE, TRUE };
enum data_source { NONE, QUERY, TARGET, BOTH };
enum algo_type { UNKNOWN, GLOBAL, SEMI_GLOBAL, LOCAL, MICROLOCAL, BANDED, KSW };
enum operation_on_seq { FORWARD_NATURAL, REVERSE_NATURAL, FORWARD_COMPLEMENT, REVERSE_COMPLEMENT };

/* ---- host staging page: src/gasal.h:75-83 --------------------------------------------------- */
struct host_batch {
  uint8_t* data;
  uint32_t page_size;
  uint32_t data_size;
  uint32_t offset;
  int is_locked;
  struct host_batch* next;
};
typedef struct host_batch host_batch_t;

/* ---- results, SoA: src/gasal.h:86-95 -------------------------------------------------------- */
struct gasal_res {
  int32_t* aln_score;
  int32_t* query_batch_end;
  int32_t* target_batch_end;
  int32_t* query_batch_start;
  int32_t* target_batch_start;
  uint8_t* cigar;          /* (count<<2)|op bytes, REVERSE alignment order, at host_query_batch_offsets[i] */
  uint32_t* n_cigar_ops;
};
typedef struct gasal_res gasal_res_t;

/* ---- per-stream storage: src/gasal.h:98-150 (same fields, same order; one private tail field) - */
typedef struct {
  uint8_t* unpacked_query_batch;
  uint8_t* unpacked_target_batch;
  uint32_t* packed_query_batch;
  uint32_t* packed_target_batch;
  uint32_t* query_batch_offsets;
  uint32_t* target_batch_offsets;
  uint32_t* query_batch_lens;
  uint32_t* target_batch_lens;

  uint32_t* host_seed_scores;
  uint32_t* seed_scores;

  host_batch_t* extensible_host_unpacked_query_batch;
  host_batch_t* extensible_host_unpacked_target_batch;

  uint8_t* host_query_op;
  uint8_t* host_target_op;
  uint8_t* query_op;
  uint8_t* target_op;

  uint32_t* host_query_batch_offsets;
  uint32_t* host_target_batch_offsets;
  uint32_t* host_query_batch_lens;
  uint32_t* host_target_batch_lens;

  gasal_res_t* host_res;          /* struct on host, arrays in pinned host memory */
  gasal_res_t* device_cpy;        /* struct on host, arrays in device memory */
  gasal_res_t* device_res;        /* struct on device (kept for source compatibility; kernels take raw arrays) */

  gasal_res_t* host_res_second;
  gasal_res_t* device_res_second;
  gasal_res_t* device_cpy_second;

  uint32_t gpu_max_query_batch_bytes;
  uint32_t gpu_max_target_batch_bytes;

  uint32_t host_max_query_batch_bytes;
  uint32_t host_max_target_batch_bytes;

  uint32_t gpu_max_n_alns;
  uint32_t host_max_n_alns;
  uint32_t current_n_alns;

  uint64_t packed_tb_matrix_size;
  uint4* packed_tb_matrices;

  cudaStream_t str;
  int is_free;
  int id;

  void* b200_priv;                /* engine-private scratch (column state, events, counters); never touch */
} gasal_gpu_storage_t;

/* vector of storages: src/gasal.h:153-156 */
typedef struct {
  int n;
  gasal_gpu_storage_t* a;
} gasal_gpu_storage_v;

/* match / mismatch / gap penalties, all positive: src/gasal.h:160-165 */
typedef struct {
  int32_t match;
  int32_t mismatch;
  int32_t gap_open;
  int32_t gap_extend;
} gasal_subst_scores;

/* ---- run-time options the library consumes: the POD image of `class Parameters` ---------------
 * (src/args_parser.h:24-68; defaults src/args_parser.cpp:12-29).  Field names are the reference's. */
typedef struct {
  int32_t algo;                      /* enum algo_type */
  int32_t start_pos;                 /* enum comp_start */
  int32_t secondBest;                /* enum Bool */
  int32_t isPacked;                  /* input already 4-bit packed (src/ctors.cpp:64-67) */
  int32_t isReverseComplement;       /* apply host_query_op / host_target_op */
  int32_t semiglobal_skipping_head;  /* enum data_source */
  int32_t semiglobal_skipping_tail;  /* enum data_source */
  int32_t k_band;
} gasal_params_t;

/* defaults of Parameters::Parameters (algo is left UNKNOWN there; callers set it) */
void gasal_params_default(gasal_params_t* p);

/* ---- src/ctors.h:4-16 ------------------------------------------------------------------------ */
gasal_gpu_storage_v gasal_init_gpu_storage_v(int n_streams);
void gasal_init_streams(gasal_gpu_storage_v* gpu_storage_vec, int max_query_len, int max_target_len, int max_n_alns,
                        const gasal_params_t* params);
void gasal_destroy_streams(gasal_gpu_storage_v* gpu_storage_vec, const gasal_params_t* params);
void gasal_destroy_gpu_storage_v(gasal_gpu_storage_v* gpu_storage_vec);
/* Deprecated in the reference (src/ctors.h:9,11, ctors.cpp:202-252) but still declared there: (re)allocate / free only the DEVICE
 * buffers of one storage (sequences, lens, offsets, device_res from an existing device_cpy).  Kept so that old callers link. */
void gasal_gpu_mem_alloc(gasal_gpu_storage_t* gpu_storage, int gpu_max_query_batch_bytes, int gpu_max_target_batch_bytes, int gpu_max_n_alns,
                         const gasal_params_t* params);
void gasal_gpu_mem_free(gasal_gpu_storage_t* gpu_storage, const gasal_params_t* params);

/* ---- src/gasal_align.h:114-120 --------------------------------------------------------------- */
void gasal_copy_subst_scores(gasal_subst_scores* subst);
/* Enqueue H2D copies, pack (+ reverse/complement), the DP kernel(s) (+ traceback) and D2H copies on
 * gpu_storage->str and return without synchronising; sets is_free = 0.  Invalid arguments or a CUDA
 * error print "[GASAL ERROR:] ..." to stderr and exit(EXIT_FAILURE), like the reference
 * (src/gasal_align.cu:29-307). */
void gasal_aln_async(gasal_gpu_storage_t* gpu_storage, const uint32_t actual_query_batch_bytes,
                     const uint32_t actual_target_batch_bytes, const uint32_t actual_n_alns, const gasal_params_t* params);
/* 0: finished (host pages reset, is_free = 1, current_n_alns = 0); -1: still running; -2: nothing launched
 * (src/gasal_align.cu:310-326) */
int gasal_is_aln_async_done(gasal_gpu_storage_t* gpu_storage);

/* ---- src/host_batch.h:9-17 ------------------------------------------------------------------- */
host_batch_t* gasal_host_batch_new(uint32_t batch_bytes, uint32_t offset);
void gasal_host_batch_destroy(host_batch_t* res);
host_batch_t* gasal_host_batch_getlast(host_batch_t* arg);
void gasal_host_batch_reset(gasal_gpu_storage_t* gpu_storage);
uint32_t gasal_host_batch_fill(gasal_gpu_storage_t* gpu_storage, uint32_t idx, const char* data, uint32_t size, enum data_source SRC);
uint32_t gasal_host_batch_add(gasal_gpu_storage_t* gpu_storage, uint32_t idx, const char* data, uint32_t size, enum data_source SRC);
uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t* gpu_storage, uint32_t idx, const char base, enum data_source SRC);
void gasal_host_batch_print(host_batch_t* res);
void gasal_host_batch_printall(host_batch_t* res);

/* ---- src/interfaces.h:8-14 ------------------------------------------------------------------- */
void gasal_host_alns_resize(gasal_gpu_storage_t* gpu_storage, int new_max_alns, const gasal_params_t* params);
void gasal_op_fill(gasal_gpu_storage_t* gpu_storage_t, uint8_t* data, uint32_t nbr_seqs_in_stream, enum data_source SRC);
/* the reference declares defaults (gpu_select = 0, isPrintingProp = true) in C++; C callers pass both */
void gasal_set_device(int gpu_select, int isPrintingProp);

/* ---- src/res.h:4-8 --------------------------------------------------------------------------- */
gasal_res_t* gasal_res_new_host(uint32_t max_n_alns, const gasal_params_t* params);
gasal_res_t* gasal_res_new_device(gasal_res_t* device_cpy);
gasal_res_t* gasal_res_new_device_cpy(uint32_t max_n_alns, const gasal_params_t* params);
void gasal_res_destroy_host(gasal_res_t* res);
void gasal_res_destroy_device(gasal_res_t* device_res, gasal_res_t* device_cpy);

/* ================================================================================================
 * Extensions (no counterpart in the reference; prefix gasalx_).
 * ============================================================================================== */

/* Run-time replacement of the reference's compile-time -DN_CODE / -DN_PENALTY (Makefile:45-49).
 * n_penalty < 0 means "N_PENALTY not defined" (N scores 0 in local/semi-global/banded/ksw). */
void gasalx_set_n_code(int n_code, int n_penalty);

/* Which DP implementation gasal_aln_async may pick:
 *   0 = automatic: s16x2 inter-task fast path when the batch qualifies (short reads), the intra-task wavefront
 *       kernel for batches with longer queries (s16x2 lanes for LOCAL when the scores fit and there are enough
 *       row-blocks, int32 lanes otherwise), else the general one-thread-per-pair int32 kernels;
 *   1 = general int32 kernels only;
 *   2 = wavefront kernel wherever it is eligible (any length), s16x2 / int32 chosen as in 0;
 *   3 = as 2 but int32 lanes only;   4 = as 2 but s16x2 lanes whenever the scores fit (testing). */
void gasalx_set_kernel_policy(int policy);

/* Semantics of the "reverse" bit of host_query_op / host_target_op.  The reference kernel
 * (src/kernels/pack_rc_seqs.h:56-212) shifts 32-bit words by 32, which is undefined in C:
 *   0 (default) = nibble reversal of the whole padded word array (what PTX shift semantics give);
 *   1           = bit-exact emulation of the reference as built by nvcc 12.9 for sm_100, where the undefined
 *                 operand is folded to 0 (first half of the words zeroed, second half = reversed first half). */
void gasalx_set_revcomp_mode(int mode);

/* Opt-in "fixed coordinates" mode (default 0 = bit-exact with the reference, which is what every parity test uses).  With 1,
 * gasal_is_aln_async_done repairs two reference quirks in the host results when a batch completes:
 *   LOCAL + WITH_START: target_batch_start is un-mirrored (the reference reports 16*floor(ts/8) + 14 - ts because
 *       src/kernels/local_kernel_template.h:41 walks `gidx + (m-1)` backwards; LOCAL + WITH_TB already reports the true start);
 *   SEMI_GLOBAL with head = tail = TARGET: query_batch_end = query length - 1 and (WITH_START) query_batch_start = 0 -- the whole
 *       query is aligned in that mode, while the reference reports ref_len and read_len - 1 - ref_len
 *       (src/kernels/semiglobal_kernel_template.h:63-64,218,381).
 * Textbook global / semi-global BOUNDARIES (the reference shifts the top row by one, global.h:57-60) are available through
 * include/nvbio_b200.h, whose Gotoh aligners use the textbook recurrence. */
void gasalx_set_coordinate_fix(int on);

/* Host-side 4-bit packing (default 1).  The reference copies 1 byte per base to the GPU and packs there (gasal_align.cu:152-201).
 * With 1, gasal_host_batch_fill / _add / _addbase pack WHILE they copy the caller's ASCII into the pinned page (AVX2), so the page,
 * the H2D copy and the device hold 4-bit words only: half the PCIe bytes, no device ASCII buffers, no pack kernel in the pipeline.
 * The caller-visible contract is unchanged (ASCII in, byte offsets and lengths in bases); host_batch_t::page_size / data_size /
 * offset stay in bases, only the private `data` bytes differ, and such a storage keeps ONE page per side that grows by
 * reallocation.  0 = the reference's layout (ASCII pages, linked 2x growth, k_pack on the GPU).  The setting is latched per storage
 * by gasal_init_streams; storages with params->isPacked always take the caller's words as they are. */
void gasalx_set_host_packing(int on);
int gasalx_storage_host_packed(const gasal_gpu_storage_t* gpu_storage);
/* 1: allocate the sequence pages as write-combined pinned memory (cudaHostAllocWriteCombined): faster for the DMA engine to read on
 * some hosts, slow for the CPU to read back.  Default 0. */
void gasalx_set_host_write_combined(int on);
/* The s16x2 fast path files every pair with a base outside A/C/G/T for its N-aware / general passes (a classification kernel reads
 * the packed sequences once).  With host packing the packer already knows whether a batch holds any such base, and a clean batch
 * skips those launches.  1 = run them regardless (testing); default 0. */
void gasalx_set_fast_path_passes(int always);

/* Device-side timing of the last finished batch of a storage, in milliseconds, measured with CUDA
 * events on gpu_storage->str: [0] whole pipeline (first H2D .. last D2H), [1] pack(+revcomp),
 * [2] DP kernels (+traceback).  Returns 0 on success, -1 if the batch has not finished. */
int gasalx_last_batch_ms(gasal_gpu_storage_t* gpu_storage, float out_ms[3]);

/* The survey's primary time over a GROUP of storages of one device whose batches were in flight together: milliseconds from the
 * earliest "first H2D enqueued" event to the latest "last D2H complete" event (CUDA events on the storages' streams).  Returns 0,
 * or -1 if one of the batches has not finished. */
int gasalx_span_ms(gasal_gpu_storage_t* const* storages, int n, float* out_ms);

/* Number of kernels the engine has launched on this storage since init (for launch accounting). */
uint64_t gasalx_kernel_launches(const gasal_gpu_storage_t* gpu_storage);

/* Name of the DP kernel family used by the last batch of this storage ("s16x2", "int32", ...). */
const char* gasalx_last_kernel_name(const gasal_gpu_storage_t* gpu_storage);

/* Device-resident entry: run pack -> DP -> (traceback) for a batch whose UNPACKED (or, with
 * params->isPacked, packed) sequences, lens and offsets are ALREADY in the storage's device buffers
 * (no host<->device copies).  Used to time the kernels alone.  Results stay in device_cpy arrays. */
void gasalx_aln_device_resident(gasal_gpu_storage_t* gpu_storage, uint32_t actual_query_batch_bytes,
                                uint32_t actual_target_batch_bytes, uint32_t actual_n_alns, const gasal_params_t* params);

/* Benchmark helper: mark the storage's pinned host pages as holding `q_bytes` / `t_bytes` of staged
 * sequence data again after gasal_is_aln_async_done reset them, WITHOUT re-copying (the bytes are
 * still there).  Lets a caller re-submit the same pre-filled pinned batch. */
void gasalx_host_batch_rearm(gasal_gpu_storage_t* gpu_storage, uint32_t q_bytes, uint32_t t_bytes, uint32_t n_alns);

/* ---- host batcher / multi-GPU sharder (genomics_gpu_b200/csrc/batcher.cpp) -------------------------
 * The reference leaves batching to the caller and has no multi-GPU mode; these entry points take ALL pairs
 * of a job, length-bin them, cut batches, spread the batches over the GPUs by LPT on DP cells, stream them
 * through gasal_aln_async on per-storage streams (one host worker thread per GPU) and return the results
 * in the caller's original pair order. */
typedef struct {
  int32_t n_gpus;                /* GPUs of this process to use (>= 1) */
  const int32_t* gpu_ids;        /* NULL: 0 .. n_gpus-1 */
  int32_t streams_per_gpu;       /* storages (streams) per GPU; 0 = automatic: 16, never more than the GPU has batches (the reference's NB_STREAMS is 2, test_prog.cpp:12) */
  uint32_t max_batch_pairs;      /* 0 = automatic: two rounds of batches per storage, 16384 .. 262144 pairs (262144 = the reference's commented GPU-filling STREAM_BATCH_SIZE, test_prog.cpp:14) */
  uint64_t max_batch_bytes;      /* per sequence side, default 1 GiB (must stay < 4 GiB: 32-bit offsets of the GASAL2 API) */
  uint64_t max_batch_cells;      /* default 2^36 DP cells */
  int32_t keep_input_order;      /* 1: no length binning (batches follow input order, like test_prog) */
  int32_t host_threads_per_gpu;  /* host threads that stage batches / gather results per GPU; 0 = automatic: half the cores divided by the GPUs, 2 .. 8 */
} gasalx_sched_t;

typedef struct {                 /* any pointer may be NULL; arrays have one entry per pair of the job */
  int32_t *aln_score, *query_batch_end, *target_batch_end, *query_batch_start, *target_batch_start;
  int32_t *aln_score2, *query_batch_end2, *target_batch_end2;
  uint8_t* cigar;                /* WITH_TB: same layout as the caller's query buffer (pair i writes at q_offs[i]) */
  uint32_t* n_cigar_ops;
} gasalx_results_t;

void gasalx_sched_default(gasalx_sched_t* s);
/* pure host planning (no CUDA): number of batches; optional outputs order[n], batch_start[nb+1], batch_gpu[nb], batch_cells[nb] (first `cap` batches) */
int64_t gasalx_plan_batches(const uint32_t* q_lens, const uint32_t* t_lens, uint64_t n, const gasalx_sched_t* sched,
                            uint32_t* order, uint32_t* batch_start, int32_t* batch_gpu, uint64_t* batch_cells, uint64_t cap);
/* align all pairs; sequences concatenated in gasal_host_batch_fill layout at 64-bit byte offsets. 0 on success.
 * gpu_ms (optional, n_gpus doubles): device time per GPU from CUDA events. */
int gasalx_align_all(const uint8_t* q_bytes, const uint8_t* t_bytes, const uint64_t* q_offs, const uint64_t* t_offs,
                     const uint32_t* q_lens, const uint32_t* t_lens, const uint8_t* q_op, const uint8_t* t_op, const uint32_t* seed_scores,
                     uint64_t n, const gasal_params_t* params, const gasalx_sched_t* sched, gasalx_results_t* out, double* gpu_ms);

/* gasalx_align_all keeps one set of storages per device between calls (pinned pages are expensive to create); this frees them. */
void gasalx_batcher_release(void);

/* ---- caller-side I/O (replaces the file loop and CIGAR printer of test_prog.cpp:79-137,382-428) ---- */
typedef struct {
  uint64_t n;                          /* pairs */
  uint8_t* q; uint8_t* t;              /* concatenated sequences, each padded with N_CODE to a multiple of 8 (gasal_host_batch_fill layout) */
  uint64_t q_bytes, t_bytes;
  uint64_t* q_offs; uint64_t* t_offs;  /* byte offsets (multiples of 8) */
  uint32_t* q_lens; uint32_t* t_lens;  /* true lengths */
  uint8_t* q_op; uint8_t* t_op;        /* 0 '>' forward, 1 '<' reverse, 2 '/' complement, 3 '+' reverse-complement (test_prog.cpp:80-89) */
  char* q_names; char* t_names;        /* header texts, one per line */
  uint32_t max_query_len, max_target_len;
  int any_op;                          /* some sequence asks for reverse / complement: set isReverseComplement */
} gasalx_fasta_pairs_t;
/* Read a query and a target FASTA file (plain or gzip) with the same number of records. 0 on success. */
int gasalx_read_fasta_pairs(const char* query_path, const char* target_path, int n_code, gasalx_fasta_pairs_t* out);
void gasalx_free_fasta_pairs(gasalx_fasta_pairs_t* p);
/* 4-bit packing on the host, device layout (pack_rc_seqs.h:13-53): n_bytes must be a multiple of 8; words[n_bytes / 8]. */
void gasalx_pack_host(const uint8_t* ascii, uint64_t n_bytes, uint32_t* words);
/* same, for `n` bases written at BASE position `pos` of an existing word array (any alignment; neighbouring nibbles are kept) */
void gasalx_pack_host_at(uint32_t* words, uint64_t pos, const uint8_t* ascii, uint64_t n);
/* gasal_res_t::cigar bytes of one pair (reverse order) -> "12M1X3D..." as test_prog.cpp:382-428 prints it.
 * Returns the length, or -(needed capacity) if cap is too small. */
int gasalx_cigar_string(const uint8_t* cigar, uint32_t n_ops, char* out, uint32_t cap);

/* CIGAR TEXT formatted on the device (one thread per pair), for the WITH_TB batch last submitted on this storage.  Call it
 * right after gasal_aln_async(..., WITH_TB): the kernel and the two D2H copies are queued on gpu_storage->str, so host_text /
 * host_lens are valid once gasal_is_aln_async_done returns 0 (use pinned memory for a truly asynchronous copy).  The text of
 * pair p is at host_text[p * stride], NUL-terminated; host_lens[p] = its length, or 0x80000000 | needed length if it was
 * truncated to stride - 1 characters.
 *   style 0: the letters test_prog.cpp:382-428 prints (M match, X mismatch, D, I), identical to gasalx_cigar_string;
 *   style 1: SAM CIGAR: M and X merge into M; for LOCAL / BANDED the unaligned query ends become soft clips.
 * Returns 0, or -1 if the storage holds no traceback results. */
int gasalx_cigar_text_async(gasal_gpu_storage_t* gpu_storage, int style, char* host_text, uint32_t stride, uint32_t* host_lens);

/* One SAM alignment line (no trailing newline) from the fields a GASAL2 caller has after a LOCAL WITH_TB batch:
 * pos1 = 1-based leftmost target coordinate (target_batch_start + 1 + the window's offset in the reference), cigar = style-1
 * text; seq / qual may be NULL ("*").  score -> AS:i, score2 >= 0 -> XS:i (second best).  Returns the length, or
 * -(needed capacity) if cap is too small. */
int gasalx_sam_record(char* out, uint32_t cap, const char* qname, int flag, const char* rname, int64_t pos1, int mapq,
                      const char* cigar, const char* seq, uint32_t seq_len, const char* qual, int32_t score, int32_t score2);

/* Opt-in NON-FATAL error mode.  The reference prints to stderr and exit()s on invalid arguments and CUDA errors (src/gasal.h:15-22,
 * src/gasal_align.cu:32-67), and so do the gasal_* entry points above.  These two run the same code but return a status instead of
 * ending the process (a server should not die on an oversized batch): GASALX_OK, or one of the codes below with the message that
 * was printed kept in gasalx_last_error() (thread-local).  After GASALX_E_ARG nothing was queued and the storage is unchanged;
 * after GASALX_E_NOMEM / GASALX_E_CUDA the storage must be destroyed (CUDA errors are sticky). */
enum { GASALX_OK = 0, GASALX_E_ARG = 1, GASALX_E_NOMEM = 2, GASALX_E_CUDA = 3 };
int gasalx_try_aln_async(gasal_gpu_storage_t* gpu_storage, uint32_t actual_query_batch_bytes, uint32_t actual_target_batch_bytes,
                         uint32_t actual_n_alns, const gasal_params_t* params);
int gasalx_try_init_streams(gasal_gpu_storage_v* gpu_storage_vec, int max_query_len, int max_target_len, int max_n_alns, const gasal_params_t* params);
const char* gasalx_last_error(void);

/* Version string of the engine. */
const char* gasalx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GASAL_B200_H */
