/* oracle/nvbio_compat.c -- TEST INFRASTRUCTURE ONLY (the product never links or calls this).
 *
 * CPU restatement of the scoring semantics of the suite's nvbio alignment module
 * (/root/reference/Non-CDP/NvB/nvbio/alignment/), the arithmetic behind
 * aln::BatchedAlignmentScore (batched.h:313-335) and sw-benchmark (NvB/sw-benchmark/sw-benchmark.cu:366-444):
 *
 *   pattern = horizontal string, index j = 1..M;  text = vertical string, index i = 1..N
 *   (gotoh/gotoh_inl.h:473-593, PatternBlockingTag)
 *     F(i,j) = max( F(i-1,j) + Ge_t , H(i-1,j) + Go_t )        gap in the pattern (consumes text)
 *     E(i,j) = max( E(i,j-1) + Ge_p , H(i,j-1) + Go_p )        gap in the text (consumes pattern)
 *     H(i,j) = max( E, F, H(i-1,j-1) + S(i,j) )  [, 0 for LOCAL]
 *   a gap of length k costs Go + (k-1) Ge  (gotoh_inl.h:513-521: open and extension are alternatives).
 *   boundaries (gotoh_inl.h:77-88 column init, :690-695 band init):
 *     H(0,0) = 0; H(0,j) = LOCAL ? 0 : Go_p + Ge_p (j-1);  H(i,0) = GLOBAL ? Go_t + Ge_t (i-1) : 0;
 *     E(i,0) = LOCAL ? 0 : -inf;  F(0,j) = -inf.
 *   result (BestSink, sink_inl.h:52-60: `score <= new` replaces, so the LAST reported maximum wins):
 *     LOCAL       every cell, reported band by band (bands of BAND columns), row by row, column by column
 *                 (gotoh_inl.h:567-584) -> winner = maximal cell with the largest (band, i, j)
 *     SEMI_GLOBAL H(i,M) for every text row i (utils_inl.h:279-299) -> largest i among the maxima
 *     GLOBAL      H(N,M) (gotoh_inl.h:896-897)
 *     sink = (i, j), 1-based end coordinates (text, pattern).
 *   Smith-Waterman (linear gaps; sw/sw_inl.h): the same recurrences with Go = Ge = deletion (text-consuming) /
 *   insertion (pattern-consuming).  Edit distance (ed/ed_utils.h:45-52): SW with 0 / -1 / -1 / -1.
 *
 * Parity: PINNED to the reference's own host build of these headers (oracle/ref_nvbio/nvbio_ref.cpp ->
 * oracle/_ref/libnvbio_ref.so) by tests/test_nvbio_oracle.py and to the golden vectors generated from it
 * (tests/golden/nvbio_scores.json, generator tools/make_nvbio_golden.py).  Domain: every H/E/F value inside the
 * reference's int16 column storage (gotoh_inl.h:599-602 asserts), min_score = INT32_MIN (no early exit).
 */
#include <stdint.h>
#include <stdlib.h>

#define NEG (-(1 << 29))
static inline int max2(int a, int b) { return a > b ? a : b; }

/* algo: 0 Gotoh (match, mismatch, gap_open, gap_ext), 1 SW (match, mismatch, deletion, insertion), 2 ED.
 * type: 0 GLOBAL, 1 LOCAL, 2 SEMI_GLOBAL.  band: the reference's BAND_LEN for the tie order (8 Gotoh, see tests).
 * out[0] score, out[1] text end i, out[2] pattern end j (1-based; -1/-1 when nothing was reported). */
int nvbio_compat_score(int algo, int type, const int32_t* scheme, const uint8_t* pattern, uint32_t M,
                       const uint8_t* text, uint32_t N, uint32_t band, int32_t* out) {
    int match = scheme[0], mismatch = scheme[1];
    int go_t, ge_t, go_p, ge_p; /* _t: gap consuming text (F, vertical), _p: gap consuming pattern (E, horizontal) */
    if (algo == 0) { go_t = go_p = scheme[2]; ge_t = ge_p = scheme[3]; }
    else if (algo == 1) { go_t = ge_t = scheme[2]; go_p = ge_p = scheme[3]; }
    else { match = 0; mismatch = -1; go_t = ge_t = go_p = ge_p = -1; }
    if (band == 0) band = 8;

    int* H = (int*)malloc(sizeof(int) * (M + 1));   /* row i-1 / i, in place */
    int* F = (int*)malloc(sizeof(int) * (M + 1));
    if (!H || !F) return -1;
    H[0] = 0;
    for (uint32_t j = 1; j <= M; ++j) { H[j] = (type == 1) ? 0 : go_p + ge_p * (int)(j - 1); F[j] = NEG; }

    int best = INT32_MIN; uint32_t bi = (uint32_t)-1, bj = (uint32_t)-1; uint64_t bkey = 0; int have = 0;
    for (uint32_t i = 1; i <= N; ++i) {
        int diag = H[0];
        H[0] = (type == 0) ? go_t + ge_t * (int)(i - 1) : 0;
        int E = (type == 1) ? 0 : NEG;
        for (uint32_t j = 1; j <= M; ++j) {
            F[j] = max2(F[j] + ge_t, H[j] + go_t);
            E = max2(E + ge_p, H[j - 1] + go_p);
            int s = (text[i - 1] == pattern[j - 1]) ? match : mismatch;
            int h = max2(max2(E, F[j]), diag + s);
            if (type == 1) h = max2(h, 0);
            diag = H[j];
            H[j] = h;
            if (type == 1) {
                uint64_t key = ((uint64_t)((j - 1) / band) << 40) | ((uint64_t)i << 20) | j;
                if (!have || h > best || (h == best && key > bkey)) { best = h; bi = i; bj = j; bkey = key; have = 1; }
            }
        }
        if (type == 2 && (!have || H[M] >= best)) { best = H[M]; bi = i; bj = M; have = 1; }
    }
    if (type == 0 && N > 0) { best = H[M]; bi = N; bj = M; have = 1; }
    out[0] = best; out[1] = (int32_t)bi; out[2] = (int32_t)bj;
    free(H); free(F);
    return 0;
}

/* Batch form: job k aligns pattern pair_pattern[k] (or k) with text pair_text[k] (or k, or 0 when n_texts == 1): the
 * host-side meaning of include/nvbio_b200.h's nvbx_batch_t.  out: scores[n], sink_i[n], sink_j[n] (sinks may be NULL). */
int nvbio_compat_batch(int algo, int type, const int32_t* scheme, const uint8_t* patterns, const uint32_t* poffs,
                       const uint8_t* texts, const uint32_t* toffs, uint32_t n_texts, const uint32_t* pair_pattern,
                       const uint32_t* pair_text, uint32_t n_pairs, uint32_t band, int32_t* scores, int32_t* sink_i,
                       int32_t* sink_j) {
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < (int64_t)n_pairs; ++k) {
        const uint32_t pi = pair_pattern ? pair_pattern[k] : (uint32_t)k;
        const uint32_t ti = pair_text ? pair_text[k] : (n_texts == 1 ? 0u : (uint32_t)k);
        int32_t out[3];
        if (nvbio_compat_score(algo, type, scheme, patterns + poffs[pi], poffs[pi + 1] - poffs[pi], texts + toffs[ti],
                               toffs[ti + 1] - toffs[ti], band, out) != 0) { rc = -1; continue; }
        scores[k] = out[0];
        if (sink_i) sink_i[k] = out[1];
        if (sink_j) sink_j[k] = out[2];
    }
    return rc;
}
