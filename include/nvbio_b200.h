/* nvbio_b200.h -- C ABI of the nvbio-style batched alignment SCORING adapter.
 *
 * Replaces, for the suite's second alignment benchmark, the path
 *     aln::BatchedAlignmentScore<stream_type, DeviceThreadScheduler>::enact( stream )
 *         (Non-CDP/NvB/nvbio/alignment/batched.h:313-335, batched_inl.h; one CUDA thread per (pattern, text) job calling
 *          aln::alignment_score, alignment.h:456-530)
 * as driven by NvB/sw-benchmark/sw-benchmark.cu:340-444 (batch_score_profile_all: Gotoh global / semi-global / local and
 * edit-distance semi-global of 256 k reads against a text, BestSink scores written to an int16 array).
 *
 * The templated C++ `stream_type` (batched.h:239-300: aligner(), size(), pattern_length(i), text_length(i),
 * load_strings(i), output(i)) becomes two PODs: nvbx_aligner_t (aligner_type + scoring scheme) and nvbx_batch_t (the
 * strings of every job).  nvbx_stream_enact is BatchedAlignmentScore::enact; nvbx_stream_results is stream.output(i)
 * for a BestSink<int32> (sink_inl.h:36-60).  Plain pointers and sizes only; host pointers unless stated otherwise.
 *
 * Semantics (bit-exact with the reference's host code path for every job whose DP values stay inside the reference's
 * int16 column storage; min_score defaults to INT32_MIN i.e. no early exit -- what sw-benchmark uses -- see nvbx_stream_set_min_scores):
 *   pattern = horizontal string j = 1..M, text = vertical string i = 1..N (gotoh/gotoh_inl.h:473-593)
 *   F(i,j) = max(F(i-1,j) + Ge, H(i-1,j) + Go);  E(i,j) = max(E(i,j-1) + Ge, H(i,j-1) + Go)
 *   H(i,j) = max(E, F, H(i-1,j-1) + (text[i] == pattern[j] ? match : mismatch)) [, 0 for LOCAL]
 *   (a gap of length k costs Go + (k-1) Ge; scores are SIGNED: penalties are negative numbers, utils.h:114-134)
 *   GLOBAL: H(N,M).  SEMI_GLOBAL: pattern global, text free at both ends: max_i H(i,M).  LOCAL: max over all cells.
 *   Smith-Waterman (utils.h:92-109): linear gaps, Go = Ge = deletion for text-consuming gaps, insertion for
 *   pattern-consuming gaps.  Edit distance (ed/ed_utils.h:45-52): 0 / -1 / -1 / -1, score = -distance.
 *   Sinks are the reference's BestSink coordinates: (text end i, pattern end j), 1-based, the LAST maximum in the
 *   reference's report order (bands of 8 (Gotoh) / 16 (SW, ED) pattern columns, then text row, then column).
 */
#ifndef NVBIO_B200_H
#define NVBIO_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* aln::AlignmentType, alignment_base.h:54 (same order) */
enum { NVBX_GLOBAL = 0, NVBX_LOCAL = 1, NVBX_SEMI_GLOBAL = 2 };
/* aligner families: aln::GotohAligner / SmithWatermanAligner / EditDistanceAligner (alignment_base.h) */
enum { NVBX_GOTOH = 0, NVBX_SMITH_WATERMAN = 1, NVBX_EDIT_DISTANCE = 2 };
/* how the symbol bytes are read: NVBX_CODES = small integer codes (0..3 = A,C,G,T as in nvbio's DNA alphabet; any other
 * byte value is a symbol of its own), NVBX_ASCII = raw characters.  Either way two symbols match iff their bytes are
 * equal (utils.h:124); the alphabet only tells the packer which bytes are the four bases of the 2-bit fast path. */
enum { NVBX_CODES = 0, NVBX_ASCII = 1 };

typedef struct {
  int32_t algo;        /* NVBX_GOTOH / NVBX_SMITH_WATERMAN / NVBX_EDIT_DISTANCE */
  int32_t type;        /* NVBX_GLOBAL / NVBX_LOCAL / NVBX_SEMI_GLOBAL */
  int32_t match;       /* SimpleGotohScheme / SimpleSmithWatermanScheme fields (utils.h:92-134); ignored for edit distance */
  int32_t mismatch;
  int32_t gap_open;    /* Gotoh: m_gap_open (cost of the FIRST gap symbol);  SW: m_deletion */
  int32_t gap_ext;     /* Gotoh: m_gap_ext;                                  SW: m_insertion */
} nvbx_aligner_t;

/* The jobs of one batch.  Sequences are concatenated; sequence s of a set occupies bytes [offsets[s], offsets[s+1]).
 * Job k aligns pattern pair_pattern[k] with text pair_text[k].  pair_pattern == NULL means job k uses pattern k;
 * pair_text == NULL means job k uses text k, or text 0 when n_texts == 1 (sw-benchmark's all-against-one shape). */
typedef struct {
  const uint8_t* patterns; const uint32_t* pattern_offsets; uint32_t n_patterns;   /* offsets: n_patterns + 1 entries */
  const uint8_t* texts;    const uint32_t* text_offsets;    uint32_t n_texts;      /* offsets: n_texts + 1 entries */
  const uint32_t* pair_pattern; const uint32_t* pair_text; uint32_t n_pairs;
  int32_t alphabet;                                                                /* NVBX_CODES / NVBX_ASCII */
} nvbx_batch_t;

typedef struct nvbx_stream nvbx_stream_t;    /* opaque: the device-resident batch (AlignmentStream) + scratch */

/* Upload a batch (H2D on `cuda_stream`, NULL = default stream) and 2-bit pack it on the device.  NULL on error. */
nvbx_stream_t* nvbx_stream_create(const nvbx_batch_t* batch, void* cuda_stream);
/* BatchedAlignmentScore::enact: queue the scoring of every job on the stream's CUDA stream; returns without waiting.
 * want_sinks != 0 also computes the BestSink coordinates (GLOBAL / SEMI_GLOBAL stay on the fast path, LOCAL uses the int32
 * kernel).  0, or -1 on invalid arguments. */
int nvbx_stream_enact(nvbx_stream_t* s, const nvbx_aligner_t* aligner, int want_sinks);
/* The reference's per-job `context_type::min_score` (batched.h:249; alignment.h:456-530 "threshold alignment score"): the DP of a job
 * is given up after a band of BAND_LEN pattern columns (8 Gotoh, 16 Smith-Waterman / edit distance; never after the last one) once
 * max_i H(i, last column of the band) + (columns still to come) * match < min_score (gotoh_inl.h:826-830, sw_inl.h:699-701).  A job
 * that gives up returns what its sink held: LOCAL the best cell of the bands processed so far, GLOBAL / SEMI_GLOBAL BestSink's
 * initial state (score -(1 << 30), sinks -1).  per_job: n_pairs host values, or NULL to use `uniform` for every job;
 * (NULL, INT32_MIN) -- the default, what sw-benchmark passes -- switches the threshold off.  Applies to the following
 * nvbx_stream_enact calls (full-matrix scoring), which then run on the band-of-8 int32 kernel ("int32-minscore"). */
int nvbx_stream_set_min_scores(nvbx_stream_t* s, const int32_t* per_job, int32_t uniform);
/* stream.output(i) for all i: wait for the last enact and copy scores (and sinks, if they were computed and the
 * pointers are non-NULL) to host arrays of n_pairs int32. */
int nvbx_stream_results(nvbx_stream_t* s, int32_t* scores, int32_t* sink_text, int32_t* sink_pattern);
/* device pointer to the int32 scores of the last enact (valid until the next enact / destroy) */
const int32_t* nvbx_stream_scores_device(nvbx_stream_t* s);
float nvbx_stream_last_ms(nvbx_stream_t* s);            /* CUDA-event time of the last enact (synchronises) */
uint64_t nvbx_stream_cells(const nvbx_stream_t* s);     /* sum over jobs of M * N (sw-benchmark.cu:380 GCUPS numerator) */
const char* nvbx_stream_kernel(const nvbx_stream_t* s); /* last enact: "s16x2-strip", "s16x2", "int32" (ordered costs), "int32-any", "banded-int32", "int32+tb", "banded-int32+tb" */
void nvbx_stream_destroy(nvbx_stream_t* s);

/* ---- banded scoring and traceback batches (round 2) ------------------------------------------------------------------------------
 * BatchedBandedAlignmentScore<BAND_LEN, stream_type>::enact (batched.h:337-357; banded_alignment_score<BAND_LEN>,
 * gotoh/gotoh_banded_inl.h:405-627, sw/sw_banded_inl.h:341-512, ed/ed_banded_inl.h): row i = pattern symbol i, band entry j =
 * text symbol i + j, j < band_len.  band_len is one of the reference's instantiations 3, 5, 7, 15, 31
 * (nvBowtie score_*_inl.h, alignment_test.cu).  Results through nvbx_stream_results: sinks are (text end, pattern end), 1-based.
 * Reference behaviour that is reproduced: a job whose text is shorter than its pattern reports nothing (score -(1 << 30) = the reference's Field_traits<int32>::min(), sinks -1);
 * GLOBAL reports band entry band_len - 1 of the last row, i.e. text end M + band_len - 1, whatever the text length; LOCAL reports
 * every band cell, also those past the end of the text; band_len 31 uses the reference's 2-bit text cache (cached text symbols
 * are truncated to their low two bits: meant for 2-bit DNA codes); the banded Smith-Waterman prices the pattern-consuming move
 * with `gap_open` (m_deletion) and the text-consuming one with `gap_ext` (m_insertion) -- the opposite of the full-matrix code.
 * Domain: text length >= band_len - 1 (the reference reads the first band_len - 1 text symbols unguarded; here positions past
 * the end read as 255).  0, or -1 on invalid arguments. */
int nvbx_stream_enact_banded(nvbx_stream_t* s, const nvbx_aligner_t* aligner, int band_len, int want_sinks);

/* An alignment as the reference's traceback returns it: aln::Alignment<int32> (alignment_base.h:124-136: score, source and sink
 * cells) plus what its Backtracer model receives (alignment.h:138-150): the soft-clip lengths and the edit operations. */
typedef struct {
  int32_t score;
  int32_t source_text, source_pattern;   /* where the walk ended: 0-based START offsets (text, pattern); -1 / -1: nothing reported */
  int32_t sink_text, sink_pattern;       /* BestSink: 1-based END coordinates (text, pattern) */
  uint32_t n_ops;                        /* number of edit operations pushed */
  uint32_t clip_end, clip_begin;         /* backtracer.clip(): pattern symbols after the alignment / before it */
} nvbx_alignment_t;
enum { NVBX_OP_SUBSTITUTION = 0, NVBX_OP_INSERTION = 1, NVBX_OP_DELETION = 2 };   /* aln::DirectionVector values a Backtracer sees */

/* BatchedAlignmentTraceback<CHECKPOINTS, stream_type>::enact (batched.h:436-451; alignment_traceback, alignment_inl.h:361-467)
 * when band_len == 0, BatchedBandedAlignmentTraceback<BAND_LEN, CHECKPOINTS, stream_type>::enact (batched.h:463-478;
 * banded_alignment_traceback, banded_inl.h:363-427) when band_len is 3, 5, 7, 15 or 31: score, BestSink, source cell and the
 * edit operations of every job, queued on the stream's CUDA stream.  The reference recomputes 4-bit flow windows from short2
 * checkpoints; here the whole flow matrix of a chunk of jobs is kept in HBM (0.5 B per cell) and walked by a second kernel.
 * Banded domain: text length >= pattern length + band_len - 1 (the reference's window recomputation reads past the text
 * otherwise: undefined there). */
int nvbx_stream_traceback(nvbx_stream_t* s, const nvbx_aligner_t* aligner, int band_len);
/* bytes reserved per job in the edit-op buffer (max pattern + max text + 64, rounded up to 4) */
uint32_t nvbx_stream_ops_stride(const nvbx_stream_t* s);
/* wait for the last nvbx_stream_traceback; alns: n_pairs entries; ops (may be NULL): n_pairs * nvbx_stream_ops_stride(s) bytes,
 * job k's operations at ops[k * stride .. + n_ops) in PUSH order -- from the END of the alignment to its source, as the
 * reference's backtracer receives them (alignment.h:151-153). */
int nvbx_stream_alignments(nvbx_stream_t* s, nvbx_alignment_t* alns, uint8_t* ops);

/* One-call form with host buffers: create, enact, results, destroy.  ms[0..2] (optional) = upload + pack, scoring,
 * download in milliseconds.  sink pointers may be NULL (then the score-only fast path is eligible). */
int nvbx_batched_score_host(const nvbx_aligner_t* aligner, const nvbx_batch_t* batch, int32_t* scores, int32_t* sink_text,
                            int32_t* sink_pattern, float ms[3]);

/* 0 = automatic: the quad-strip s16x2 kernel ("s16x2-strip": one job per thread, column state in shared memory, nothing parked in
 *     HBM) for scores and GLOBAL / SEMI_GLOBAL sinks of patterns up to 352 symbols, the band-blocked s16x2 kernel ("s16x2": two jobs
 *     per thread, the reference's blocking) where its range proof allows it, int32 otherwise;
 * 1 = always the int32 kernel, 2 = never the quad-strip kernel (A/B testing and parity tests). */
void nvbx_set_policy(int policy);
uint64_t nvbx_kernel_launches(void);
const char* nvbx_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NVBIO_B200_H */
