// oracle/ref_nvbio/nvbio_ref.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin C entry point over the reference's OWN host code path for nvbio-style scoring:
//   nvbio::aln::alignment_score<MAX_TEXT_LEN>( aligner, pattern, quals, text, min_score, sink )
//   (/root/reference/Non-CDP/NvB/nvbio/alignment/alignment.h:511-530, gotoh/gotoh_inl.h, sw/sw_inl.h, ed/ed_inl.h)
// compiled with g++ from the headers where they lie (recipe: oracle/Makefile, target ref_nvbio); nothing is copied.
// Used to pin oracle/nvbio_compat.c and to generate tests/golden/nvbio_*.json.
#include <nvbio/basic/types.h>
#include <nvbio/alignment/alignment.h>
#include <cstdint>

using namespace nvbio;

namespace {
const uint32 MAX_TEXT = 16384;
typedef vector_view<const uint8*> str_t;

template <typename aligner_t>
int run(const aligner_t aligner, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out) {
    aln::BestSink<int32> sink;
    aln::alignment_score<MAX_TEXT>(aligner, str_t(m, p), aln::trivial_quality_string(), str_t(n, t),
                                   Field_traits<int32>::min(), sink);
    out[0] = sink.score;
    out[1] = int32_t(sink.sink.x);
    out[2] = int32_t(sink.sink.y);
    return 0;
}

template <aln::AlignmentType TYPE>
int run_type(int algo, const int32_t* sc, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int32_t* out) {
    if (algo == 0) return run(aln::make_gotoh_aligner<TYPE>(aln::SimpleGotohScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out);
    if (algo == 1) return run(aln::make_smith_waterman_aligner<TYPE>(aln::SimpleSmithWatermanScheme(sc[0], sc[1], sc[2], sc[3])), p, m, t, n, out);
    if (algo == 2) return run(aln::make_edit_distance_aligner<TYPE>(), p, m, t, n, out);
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
