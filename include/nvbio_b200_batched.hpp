// nvbio_b200_batched.hpp -- header-only C++ front end over the C ABI of include/nvbio_b200.h, with the NAMES of the reference's
// batched scoring interface so that code written against Non-CDP/NvB/nvbio/alignment reads the same:
//
//   reference (NvB/sw-benchmark/sw-benchmark.cu:366-383, nvbio/alignment/batched.h:313-335)     here
//   -------------------------------------------------------------------------------------     ---------------------------------
//   aln::SimpleGotohScheme, aln::SimpleSmithWatermanScheme (utils.h:92-134)                     nvbio_b200::aln::  same names/fields
//   aln::make_gotoh_aligner<aln::LOCAL>( scheme )            (alignment_base.h)                 same
//   aln::make_smith_waterman_aligner<TYPE>( scheme ), aln::make_edit_distance_aligner<TYPE>()   same
//   typedef aln::BatchedAlignmentScore<stream_type, scheduler> batch_type;                      same (the scheduler tag is accepted and ignored)
//   batch_type::max_temp_storage( max_pattern_len, max_text_len, size )                         same (returns 0: scratch is internal)
//   batch.enact( stream, temp_size, temp )                                                      same
//   aln::BatchedBandedAlignmentScore<BAND_LEN, stream, scheduler>  (batched.h:337-357)           same
//   aln::BatchedAlignmentTraceback<CHECKPOINTS, stream, scheduler> (batched.h:436-451)           same, over aln::HostTracebackStream
//   aln::BatchedBandedAlignmentTraceback<BAND_LEN, CHECKPOINTS, stream, scheduler> (:463-478)    same
//
// The reference's stream_type is a user-written template over packed device strings (batched.h:239-300); here it is the concrete
// aln::HostStringStream<aligner_type>: host string sets (concatenated symbols + offsets), an optional pair list, and the output
// arrays that the reference's stream.output(i) would fill from a BestSink<int32> (score, and optionally sink.x / sink.y).
// No CPU fallback: enact() throws std::runtime_error when the library reports an error (e.g. no CUDA device).
#ifndef NVBIO_B200_BATCHED_HPP
#define NVBIO_B200_BATCHED_HPP
#include <cstdint>
#include <stdexcept>
#include <string>
#include "nvbio_b200.h"

namespace nvbio_b200 {
namespace aln {

enum AlignmentType { GLOBAL = NVBX_GLOBAL, LOCAL = NVBX_LOCAL, SEMI_GLOBAL = NVBX_SEMI_GLOBAL };   // alignment_base.h:54

struct SimpleGotohScheme {            // utils.h:114-134
  SimpleGotohScheme() {}
  SimpleGotohScheme(int32_t match, int32_t mm, int32_t gap_open, int32_t gap_ext) : m_match(match), m_mismatch(mm), m_gap_open(gap_open), m_gap_ext(gap_ext) {}
  int32_t m_match = 0, m_mismatch = 0, m_gap_open = 0, m_gap_ext = 0;
};
struct SimpleSmithWatermanScheme {    // utils.h:92-109
  SimpleSmithWatermanScheme() {}
  SimpleSmithWatermanScheme(int32_t match, int32_t mm, int32_t del, int32_t ins) : m_match(match), m_mismatch(mm), m_deletion(del), m_insertion(ins) {}
  int32_t m_match = 0, m_mismatch = 0, m_deletion = 0, m_insertion = 0;
};

// scheduler tags of batched.h:51-76: accepted for source compatibility, the engine schedules by itself
struct DeviceThreadScheduler {};
struct DeviceStagedThreadScheduler {};
struct DeviceWarpScheduler {};

template <AlignmentType T, typename scheme_type> struct GotohAligner {
  static const AlignmentType TYPE = T;
  scheme_type scheme;
  nvbx_aligner_t pod() const { nvbx_aligner_t a = {NVBX_GOTOH, (int32_t)T, scheme.m_match, scheme.m_mismatch, scheme.m_gap_open, scheme.m_gap_ext}; return a; }
};
template <AlignmentType T, typename scheme_type> struct SmithWatermanAligner {
  static const AlignmentType TYPE = T;
  scheme_type scheme;
  nvbx_aligner_t pod() const { nvbx_aligner_t a = {NVBX_SMITH_WATERMAN, (int32_t)T, scheme.m_match, scheme.m_mismatch, scheme.m_deletion, scheme.m_insertion}; return a; }
};
template <AlignmentType T> struct EditDistanceAligner {
  static const AlignmentType TYPE = T;
  nvbx_aligner_t pod() const { nvbx_aligner_t a = {NVBX_EDIT_DISTANCE, (int32_t)T, 0, -1, -1, -1}; return a; }
};
template <AlignmentType T, typename scheme_type> GotohAligner<T, scheme_type> make_gotoh_aligner(const scheme_type& s) { GotohAligner<T, scheme_type> a; a.scheme = s; return a; }
template <AlignmentType T, typename scheme_type> SmithWatermanAligner<T, scheme_type> make_smith_waterman_aligner(const scheme_type& s) { SmithWatermanAligner<T, scheme_type> a; a.scheme = s; return a; }
template <AlignmentType T> EditDistanceAligner<T> make_edit_distance_aligner() { return EditDistanceAligner<T>(); }

// The jobs of one batch and where their results go (the reference's stream_type with a BestSink<int32> per job).
template <typename aligner_t>
struct HostStringStream {
  typedef aligner_t aligner_type;
  HostStringStream(const aligner_type& aligner, uint32_t n_jobs, const uint8_t* patterns, const uint32_t* pattern_offsets, uint32_t n_patterns,
                   const uint8_t* texts, const uint32_t* text_offsets, uint32_t n_texts, int32_t* scores, int32_t* sink_text = nullptr,
                   int32_t* sink_pattern = nullptr, const uint32_t* pair_pattern = nullptr, const uint32_t* pair_text = nullptr,
                   int alphabet = NVBX_CODES)
      : m_aligner(aligner), m_scores(scores), m_sink_text(sink_text), m_sink_pattern(sink_pattern), m_max_pattern(0), m_max_text(0),
        m_min_scores(nullptr), m_min_score(INT32_MIN) {
    m_batch.patterns = patterns; m_batch.pattern_offsets = pattern_offsets; m_batch.n_patterns = n_patterns;
    m_batch.texts = texts; m_batch.text_offsets = text_offsets; m_batch.n_texts = n_texts;
    m_batch.pair_pattern = pair_pattern; m_batch.pair_text = pair_text; m_batch.n_pairs = n_jobs; m_batch.alphabet = alphabet;
    for (uint32_t i = 0; i < n_patterns; ++i) { const uint32_t l = pattern_offsets[i + 1] - pattern_offsets[i]; if (l > m_max_pattern) m_max_pattern = l; }
    for (uint32_t i = 0; i < n_texts; ++i) { const uint32_t l = text_offsets[i + 1] - text_offsets[i]; if (l > m_max_text) m_max_text = l; }
  }
  aligner_type aligner() const { return m_aligner; }
  uint32_t size() const { return m_batch.n_pairs; }
  uint32_t max_pattern_length() const { return m_max_pattern; }
  uint32_t max_text_length() const { return m_max_text; }

  aligner_type m_aligner;
  nvbx_batch_t m_batch;
  int32_t *m_scores, *m_sink_text, *m_sink_pattern;
  uint32_t m_max_pattern, m_max_text;
  // the reference's context_type::min_score (batched.h:249): per-job thresholds (n_jobs values) or one for every job; INT32_MIN = off
  void set_min_scores(const int32_t* per_job, int32_t uniform = INT32_MIN) { m_min_scores = per_job; m_min_score = uniform; }
  const int32_t* m_min_scores; int32_t m_min_score;
};

template <typename stream_type, typename algorithm_type = DeviceThreadScheduler>
struct BatchedAlignmentScore {
  typedef typename stream_type::aligner_type aligner_type;
  static uint64_t min_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  static uint64_t max_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  // enact the batch execution (batched.h:330-334); ms (optional) = upload + pack, scoring, download in milliseconds
  void enact(stream_type stream, uint64_t /*temp_size*/ = 0u, uint8_t* /*temp*/ = nullptr, float* ms = nullptr) {
    const nvbx_aligner_t a = stream.aligner().pod();
    if (stream.m_min_scores || stream.m_min_score != INT32_MIN) {        // thresholds: the stream form (create, set, enact, results)
      nvbx_stream_t* h = nvbx_stream_create(&stream.m_batch, nullptr);
      const bool sinks = stream.m_sink_text || stream.m_sink_pattern;
      const bool ok = h && nvbx_stream_set_min_scores(h, stream.m_min_scores, stream.m_min_score) == 0 && nvbx_stream_enact(h, &a, sinks ? 1 : 0) == 0 &&
                      nvbx_stream_results(h, stream.m_scores, stream.m_sink_text, stream.m_sink_pattern) == 0;
      if (h) nvbx_stream_destroy(h);
      if (!ok) throw std::runtime_error(std::string("BatchedAlignmentScore::enact: ") + nvbx_last_error());
      return;
    }
    if (nvbx_batched_score_host(&a, &stream.m_batch, stream.m_scores, stream.m_sink_text, stream.m_sink_pattern, ms) != 0)
      throw std::runtime_error(std::string("BatchedAlignmentScore::enact: ") + nvbx_last_error());
  }
};

// A stream whose jobs also want their alignments traced back: the reference's traceback stream_type hands every job a
// backtracer in its context (batched.h:358-433); here the Alignment<int32> of job i lands in alignments[i] and the operations its
// backtracer would receive in ops[i * ops_stride .. + alignments[i].n_ops), push order (end of the alignment first).
template <typename aligner_t>
struct HostTracebackStream : HostStringStream<aligner_t> {
  HostTracebackStream(const aligner_t& aligner, uint32_t n_jobs, const uint8_t* patterns, const uint32_t* pattern_offsets, uint32_t n_patterns,
                      const uint8_t* texts, const uint32_t* text_offsets, uint32_t n_texts, nvbx_alignment_t* alignments, uint8_t* ops, uint32_t ops_stride,
                      const uint32_t* pair_pattern = nullptr, const uint32_t* pair_text = nullptr, int alphabet = NVBX_CODES)
      : HostStringStream<aligner_t>(aligner, n_jobs, patterns, pattern_offsets, n_patterns, texts, text_offsets, n_texts, nullptr, nullptr, nullptr,
                                    pair_pattern, pair_text, alphabet),
        m_alignments(alignments), m_ops(ops), m_ops_stride(ops_stride) {}
  // bytes per job the op buffer needs (what nvbx_stream_ops_stride reports for this batch)
  uint32_t ops_stride() const { return ((this->m_max_pattern + this->m_max_text + 64u) + 3u) & ~3u; }
  nvbx_alignment_t* m_alignments; uint8_t* m_ops; uint32_t m_ops_stride;
};

namespace detail {
struct StreamHandle {            // nvbx_stream_t with scope-bound lifetime
  explicit StreamHandle(const nvbx_batch_t* b, const char* who) : h(nvbx_stream_create(b, nullptr)) {
    if (!h) throw std::runtime_error(std::string(who) + ": " + nvbx_last_error());
  }
  ~StreamHandle() { nvbx_stream_destroy(h); }
  nvbx_stream_t* h;
};
template <typename stream_type>
inline void traceback(stream_type& stream, int band_len, const char* who) {
  StreamHandle s(&stream.m_batch, who);
  const nvbx_aligner_t a = stream.aligner().pod();
  if (stream.m_ops && stream.m_ops_stride != nvbx_stream_ops_stride(s.h)) throw std::runtime_error(std::string(who) + ": ops_stride must be stream.ops_stride()");
  if (nvbx_stream_traceback(s.h, &a, band_len) != 0 || nvbx_stream_alignments(s.h, stream.m_alignments, stream.m_ops) != 0)
    throw std::runtime_error(std::string(who) + ": " + nvbx_last_error());
}
}  // namespace detail

// batched.h:337-357
template <uint32_t BAND_LEN, typename stream_type, typename algorithm_type = DeviceThreadScheduler>
struct BatchedBandedAlignmentScore {
  typedef typename stream_type::aligner_type aligner_type;
  static uint64_t min_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  static uint64_t max_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  void enact(stream_type stream, uint64_t /*temp_size*/ = 0u, uint8_t* /*temp*/ = nullptr) {
    detail::StreamHandle s(&stream.m_batch, "BatchedBandedAlignmentScore::enact");
    const nvbx_aligner_t a = stream.aligner().pod();
    if (nvbx_stream_enact_banded(s.h, &a, (int)BAND_LEN, (stream.m_sink_text || stream.m_sink_pattern) ? 1 : 0) != 0 ||
        nvbx_stream_results(s.h, stream.m_scores, stream.m_sink_text, stream.m_sink_pattern) != 0)
      throw std::runtime_error(std::string("BatchedBandedAlignmentScore::enact: ") + nvbx_last_error());
  }
};
// batched.h:436-451 (CHECKPOINTS is accepted and ignored: the flow matrix is kept whole)
template <uint32_t CHECKPOINTS, typename stream_type, typename algorithm_type = DeviceThreadScheduler>
struct BatchedAlignmentTraceback {
  static uint64_t min_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  static uint64_t max_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  void enact(stream_type stream, uint64_t /*temp_size*/ = 0u, uint8_t* /*temp*/ = nullptr) { detail::traceback(stream, 0, "BatchedAlignmentTraceback::enact"); }
};
// batched.h:463-478
template <uint32_t BAND_LEN, uint32_t CHECKPOINTS, typename stream_type, typename algorithm_type = DeviceThreadScheduler>
struct BatchedBandedAlignmentTraceback {
  static uint64_t min_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  static uint64_t max_temp_storage(uint32_t, uint32_t, uint32_t) { return 0; }
  void enact(stream_type stream, uint64_t /*temp_size*/ = 0u, uint8_t* /*temp*/ = nullptr) { detail::traceback(stream, (int)BAND_LEN, "BatchedBandedAlignmentTraceback::enact"); }
};

}  // namespace aln
}  // namespace nvbio_b200
#endif  // NVBIO_B200_BATCHED_HPP
