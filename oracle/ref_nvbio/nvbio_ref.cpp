// oracle/ref_nvbio/nvbio_ref.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin C entry point over the reference's OWN host code path for nvbio-style scoring:
//   nvbio::aln::alignment_score<MAX_TEXT_LEN>( aligner, pattern, quals, text, min_score, sink )
//   (/root/reference/Non-CDP/NvB/nvbio/alignment/alignment.h:511-530, gotoh/gotoh_inl.h, sw/sw_inl.h, ed/ed_inl.h)
// compiled with g++ from the headers where they lie (recipe: oracle/Makefile, target ref_nvbio); nothing is copied.
// Used to pin oracle/nvbio_compat.c and to generate tests/golden/nvbio_*.json.
#include <nvbio/basic/types.h>
#include <nvbio/alignment/alignment.h>
#include <cstdint>
#include <vector>

using namespace nvbio;

namespace {
const uint32 MAX_TEXT = 16384;
typedef vector_view<const uint8*> str_t;

template <typename aligner_t>
int run(const aligner_t aligner, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out, int32_t min_score, bool want_ok) {
    aln::BestSink<int32> sink;
    const bool ok = aln::alignment_score<MAX_TEXT>(aligner, str_t(m, p), aln::trivial_quality_string(), str_t(n, t), min_score, sink);
    out[0] = sink.score;
    out[1] = int32_t(sink.sink.x);
    out[2] = int32_t(sink.sink.y);
    if (want_ok) out[3] = ok ? 1 : 0;
    return 0;
}

template <aln::AlignmentType TYPE>
int run_type(int algo, const int32_t* sc, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out,
             int32_t min_score = Field_traits<int32>::min(), bool want_ok = false) {
    if (algo == 0) return run(aln::make_gotoh_aligner<TYPE>(aln::SimpleGotohScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out, min_score, want_ok);
    if (algo == 1) return run(aln::make_smith_waterman_aligner<TYPE>(aln::SimpleSmithWatermanScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out, min_score, want_ok);
    if (algo == 2) return run(aln::make_edit_distance_aligner<TYPE>(), p, m, t, n, out, min_score, want_ok);
    return -1;
}
}  // namespace

// algo: 0 Gotoh (match, mismatch, gap_open, gap_ext), 1 Smith-Waterman (match, mismatch, deletion, insertion),
//       2 edit distance (scheme ignored).  type: 0 GLOBAL, 1 LOCAL, 2 SEMI_GLOBAL (aln::AlignmentType order).
// out[0] = sink.score, out[1] = sink.sink.x (text end), out[2] = sink.sink.y (pattern end).
extern "C" int nvbio_ref_score(int algo, int type, const int32_t* scheme, const uint8_t* pattern, uint32_t m,
                               const uint8_t* text, uint32_t n, int32_t* out) {
    if (n > MAX_TEXT) return -2;
    switch (type) {
        case 0: return run_type<aln::GLOBAL>(algo, scheme, pattern, m, text, n, out);
        case 1: return run_type<aln::LOCAL>(algo, scheme, pattern, m, text, n, out);
        case 2: return run_type<aln::SEMI_GLOBAL>(algo, scheme, pattern, m, text, n, out);
    }
    return -1;
}

// the same call with a min_score threshold (the early exit of gotoh_inl.h:826-830 / sw_inl.h:699-701); out[3] = the function's bool result
extern "C" int nvbio_ref_score_min(int algo, int type, const int32_t* scheme, const uint8_t* pattern, uint32_t m,
                                   const uint8_t* text, uint32_t n, int32_t min_score, int32_t* out) {
    if (n > MAX_TEXT) return -2;
    switch (type) {
        case 0: return run_type<aln::GLOBAL>(algo, scheme, pattern, m, text, n, out, min_score, true);
        case 1: return run_type<aln::LOCAL>(algo, scheme, pattern, m, text, n, out, min_score, true);
        case 2: return run_type<aln::SEMI_GLOBAL>(algo, scheme, pattern, m, text, n, out, min_score, true);
    }
    return -1;
}

// ---- banded scoring and traceback (round 2: the rest of SURVEY row f4) -------------------------------------------------
//   aln::banded_alignment_score<BAND_LEN>( aligner, pattern, quals, text, min_score, sink )   (banded_inl.h:57-75)
//   aln::alignment_traceback<MAX_PATTERN_LEN, MAX_TEXT_LEN, CHECKPOINTS>( aligner, pattern, quals, text, min_score, backtracer )
//                                                                                             (alignment_inl.h:489-528)
//   aln::banded_alignment_traceback<BAND_LEN, MAX_PATTERN_LEN, CHECKPOINTS>( ... )            (banded_inl.h:450-489)
namespace {
struct RefBacktracer {        // the Backtracer model of alignment.h:138-150: ops arrive from the END of the alignment to its source
    std::vector<uint8_t> ops;
    uint32_t clips[4]; int n_clips;
    RefBacktracer() : n_clips(0) { clips[0] = clips[1] = clips[2] = clips[3] = 0; }
    void clip(const uint32 n) { if (n_clips < 4) clips[n_clips] = n; ++n_clips; }
    void push(const uint8 op) { ops.push_back(op); }
};

template <uint32 BAND, typename aligner_t>
int run_banded(const aligner_t aligner, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out) {
    aln::BestSink<int32> sink;
    const bool ok = aln::banded_alignment_score<BAND>(aligner, str_t(m, p), aln::trivial_quality_string(), str_t(n, t),
                                                     Field_traits<int32>::min(), sink);
    out[0] = sink.score; out[1] = int32_t(sink.sink.x); out[2] = int32_t(sink.sink.y); out[3] = ok ? 1 : 0;
    return 0;
}
template <typename aligner_t>
int run_banded_b(int band, const aligner_t aligner, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out) {
    switch (band) {
        case 3: return run_banded<3>(aligner, p, m, t, n, out);
        case 5: return run_banded<5>(aligner, p, m, t, n, out);
        case 7: return run_banded<7>(aligner, p, m, t, n, out);
        case 15: return run_banded<15>(aligner, p, m, t, n, out);
        case 31: return run_banded<31>(aligner, p, m, t, n, out);
    }
    return -3;
}
template <aln::AlignmentType TYPE>
int run_banded_type(int algo, int band, const int32_t* sc, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out) {
    if (algo == 0) return run_banded_b(band, aln::make_gotoh_aligner<TYPE>(aln::SimpleGotohScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out);
    if (algo == 1) return run_banded_b(band, aln::make_smith_waterman_aligner<TYPE>(aln::SimpleSmithWatermanScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out);
    if (algo == 2) return run_banded_b(band, aln::make_edit_distance_aligner<TYPE>(), p, m, t, n, out);
    return -1;
}

const uint32 TB_MAX = 1024, TB_CHECKPOINTS = 32;
template <typename aligner_t>
int run_tb(const aligner_t aligner, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out, uint8_t* ops, uint32_t cap) {
    RefBacktracer bt;
    const aln::Alignment<int32> a = aln::alignment_traceback<TB_MAX, TB_MAX, TB_CHECKPOINTS>(
        aligner, str_t(m, p), aln::trivial_quality_string(), str_t(n, t), Field_traits<int32>::min(), bt);
    out[0] = a.score; out[1] = int32_t(a.source.x); out[2] = int32_t(a.source.y); out[3] = int32_t(a.sink.x); out[4] = int32_t(a.sink.y);
    out[5] = int32_t(bt.ops.size()); out[6] = int32_t(bt.clips[0]); out[7] = int32_t(bt.clips[1]); out[8] = bt.n_clips;
    if (bt.ops.size() > cap) return -4;
    for (size_t i = 0; i < bt.ops.size(); ++i) ops[i] = bt.ops[i];
    return 0;
}
template <aln::AlignmentType TYPE>
int run_tb_type(int algo, const int32_t* sc, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out, uint8_t* ops, uint32_t cap) {
    if (algo == 0) return run_tb(aln::make_gotoh_aligner<TYPE>(aln::SimpleGotohScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out, ops, cap);
    if (algo == 1) return run_tb(aln::make_smith_waterman_aligner<TYPE>(aln::SimpleSmithWatermanScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out, ops, cap);
    if (algo == 2) return run_tb(aln::make_edit_distance_aligner<TYPE>(), p, m, t, n, out, ops, cap);
    return -1;
}

template <uint32 BAND, typename aligner_t>
int run_btb(const aligner_t aligner, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out, uint8_t* ops, uint32_t cap) {
    RefBacktracer bt;
    const aln::Alignment<int32> a = aln::banded_alignment_traceback<BAND, TB_MAX, TB_CHECKPOINTS>(
        aligner, str_t(m, p), aln::trivial_quality_string(), str_t(n, t), Field_traits<int32>::min(), bt);
    out[0] = a.score; out[1] = int32_t(a.source.x); out[2] = int32_t(a.source.y); out[3] = int32_t(a.sink.x); out[4] = int32_t(a.sink.y);
    out[5] = int32_t(bt.ops.size()); out[6] = int32_t(bt.clips[0]); out[7] = int32_t(bt.clips[1]); out[8] = bt.n_clips;
    if (bt.ops.size() > cap) return -4;
    for (size_t i = 0; i < bt.ops.size(); ++i) ops[i] = bt.ops[i];
    return 0;
}
template <typename aligner_t>
int run_btb_b(int band, const aligner_t aligner, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out, uint8_t* ops, uint32_t cap) {
    switch (band) {
        case 7: return run_btb<7>(aligner, p, m, t, n, out, ops, cap);
        case 15: return run_btb<15>(aligner, p, m, t, n, out, ops, cap);
        case 31: return run_btb<31>(aligner, p, m, t, n, out, ops, cap);
    }
    return -3;
}
template <aln::AlignmentType TYPE>
int run_btb_type(int algo, int band, const int32_t* sc, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out, uint8_t* ops, uint32_t cap) {
    if (algo == 0) return run_btb_b(band, aln::make_gotoh_aligner<TYPE>(aln::SimpleGotohScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out, ops, cap);
    if (algo == 1) return run_btb_b(band, aln::make_smith_waterman_aligner<TYPE>(aln::SimpleSmithWatermanScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out, ops, cap);
    if (algo == 2) return run_btb_b(band, aln::make_edit_distance_aligner<TYPE>(), p, m, t, n, out, ops, cap);
    return -1;
}
}  // namespace

// band: the reference's BAND_LEN template argument (3, 5, 7, 15 or 31).  out[0..2] as nvbio_ref_score, out[3] = the function's
// bool result (false: text shorter than the pattern -- nothing was reported).
extern "C" int nvbio_ref_banded_score(int algo, int type, int band, const int32_t* scheme, const uint8_t* pattern, uint32_t m,
                                      const uint8_t* text, uint32_t n, int32_t* out) {
    switch (type) {
        case 0: return run_banded_type<aln::GLOBAL>(algo, band, scheme, pattern, m, text, n, out);
        case 1: return run_banded_type<aln::LOCAL>(algo, band, scheme, pattern, m, text, n, out);
        case 2: return run_banded_type<aln::SEMI_GLOBAL>(algo, band, scheme, pattern, m, text, n, out);
    }
    return -1;
}

// out[0] score, out[1..2] source (text, pattern), out[3..4] sink (text, pattern), out[5] number of ops, out[6] first clip
// (end of the pattern), out[7] second clip (beginning), out[8] number of clip() calls; ops[] in push order (end -> source),
// values of aln::DirectionVector: 0 SUBSTITUTION, 1 INSERTION, 2 DELETION.
extern "C" int nvbio_ref_traceback(int algo, int type, const int32_t* scheme, const uint8_t* pattern, uint32_t m,
                                   const uint8_t* text, uint32_t n, int32_t* out, uint8_t* ops, uint32_t cap) {
    if (n > TB_MAX || m > TB_MAX) return -2;
    switch (type) {
        case 0: return run_tb_type<aln::GLOBAL>(algo, scheme, pattern, m, text, n, out, ops, cap);
        case 1: return run_tb_type<aln::LOCAL>(algo, scheme, pattern, m, text, n, out, ops, cap);
        case 2: return run_tb_type<aln::SEMI_GLOBAL>(algo, scheme, pattern, m, text, n, out, ops, cap);
    }
    return -1;
}

extern "C" int nvbio_ref_banded_traceback(int algo, int type, int band, const int32_t* scheme, const uint8_t* pattern, uint32_t m,
                                          const uint8_t* text, uint32_t n, int32_t* out, uint8_t* ops, uint32_t cap) {
    if (m > TB_MAX) return -2;
    switch (type) {
        case 0: return run_btb_type<aln::GLOBAL>(algo, band, scheme, pattern, m, text, n, out, ops, cap);
        case 1: return run_btb_type<aln::LOCAL>(algo, band, scheme, pattern, m, text, n, out, ops, cap);
        case 2: return run_btb_type<aln::SEMI_GLOBAL>(algo, band, scheme, pattern, m, text, n, out, ops, cap);
    }
    return -1;
}
