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
 * reference's int16 column storage (gotoh_inl.h:599-602 asserts).  nvbio_compat_score is the min_score = INT32_MIN case (no early
 * exit, what sw-benchmark passes); nvbio_compat_score_min adds the early exit.
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

/* min_score early exit (alignment.h:456-530 `min_score`; gotoh_inl.h:826-830, sw_inl.h:699-701): after each band of `band` pattern
 * columns EXCEPT the last one, the reference takes max_score = max over the text rows of H(i, last column of the band) and gives up
 * -- returns false, the sink keeps what was reported so far -- when max_score + (columns still to come) * match < min_score.
 * LOCAL reports from every band, so it keeps the best cell of the bands processed (in their report order); GLOBAL / SEMI_GLOBAL
 * only report from the last band, so their sink stays at BestSink's initial state (score Field_traits<int32>::min() = -(1 << 30),
 * sink (-1, -1); sink_inl.h:40).  Columns only depend on the columns to their left, so the bands processed before the exit hold the
 * values of the full DP: the restatement finds the exit band from the column maxima of a full DP and, for LOCAL, rescoring the
 * pattern truncated after that band.  out[3] = the function's bool result (1: ran to the end, 0: gave up).
 * Edit distance scores a match 0, so its bound is max_score itself. */
int nvbio_compat_score_min(int algo, int type, const int32_t* scheme, const uint8_t* pattern, uint32_t M, const uint8_t* text,
                           uint32_t N, uint32_t band, int32_t min_score, int32_t* out) {
    int match = scheme[0], mismatch = scheme[1];
    int go_t, ge_t, go_p, ge_p;
    if (algo == 0) { go_t = go_p = scheme[2]; ge_t = ge_p = scheme[3]; }
    else if (algo == 1) { go_t = ge_t = scheme[2]; go_p = ge_p = scheme[3]; }
    else { match = 0; mismatch = -1; go_t = ge_t = go_p = ge_p = -1; }
    if (band == 0) band = 8;
    out[3] = 1;
    const uint32_t nb = (M + band - 1) / band;                 /* bands; the last one is never checked */
    if (nb > 1 && N > 0) {
        int* H = (int*)malloc(sizeof(int) * (M + 1));
        int* F = (int*)malloc(sizeof(int) * (M + 1));
        int* cmax = (int*)malloc(sizeof(int) * (M + 1));
        if (!H || !F || !cmax) return -1;
        H[0] = 0;
        for (uint32_t j = 1; j <= M; ++j) { H[j] = (type == 1) ? 0 : go_p + ge_p * (int)(j - 1); F[j] = NEG; cmax[j] = INT32_MIN; }
        for (uint32_t i = 1; i <= N; ++i) {
            int diag = H[0];
            H[0] = (type == 0) ? go_t + ge_t * (int)(i - 1) : 0;
            int E = (type == 1) ? 0 : NEG;
            for (uint32_t j = 1; j <= M; ++j) {
                F[j] = max2(F[j] + ge_t, H[j] + go_t);
                E = max2(E + ge_p, H[j - 1] + go_p);
                const int s = (text[i - 1] == pattern[j - 1]) ? match : mismatch;
                int h = max2(max2(E, F[j]), diag + s);
                if (type == 1) h = max2(h, 0);
                diag = H[j];
                H[j] = h;
                if (h > cmax[j]) cmax[j] = h;
            }
        }
        uint32_t exit_cols = 0;
        for (uint32_t b = 0; b + 1 < nb && !exit_cols; ++b) {
            const uint32_t c = (b + 1) * band;                 /* last column of band b (always <= M here) */
            if ((int64_t)cmax[c] + (int64_t)(M - c) * match < (int64_t)min_score) exit_cols = c;
        }
        free(H); free(F); free(cmax);
        if (exit_cols) {
            out[3] = 0;
            if (type == 1) return nvbio_compat_score(algo, type, scheme, pattern, exit_cols, text, N, band, out);
            out[0] = -(1 << 30); out[1] = -1; out[2] = -1;
            return 0;
        }
    }
    return nvbio_compat_score(algo, type, scheme, pattern, M, text, N, band, out);
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


/* batch form with the early exit: min_scores == NULL -> `min_score` for every job */
int nvbio_compat_batch_min(int algo, int type, const int32_t* scheme, const uint8_t* patterns, const uint32_t* poffs,
                           const uint8_t* texts, const uint32_t* toffs, uint32_t n_texts, const uint32_t* pair_pattern,
                           const uint32_t* pair_text, uint32_t n_pairs, uint32_t band, const int32_t* min_scores, int32_t min_score,
                           int32_t* scores, int32_t* sink_i, int32_t* sink_j) {
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < (int64_t)n_pairs; ++k) {
        const uint32_t pi = pair_pattern ? pair_pattern[k] : (uint32_t)k;
        const uint32_t ti = pair_text ? pair_text[k] : (n_texts == 1 ? 0u : (uint32_t)k);
        int32_t out[4];
        if (nvbio_compat_score_min(algo, type, scheme, patterns + poffs[pi], poffs[pi + 1] - poffs[pi], texts + toffs[ti],
                                   toffs[ti + 1] - toffs[ti], band, min_scores ? min_scores[k] : min_score, out) != 0) { rc = -1; continue; }
        scores[k] = out[0];
        if (sink_i) sink_i[k] = out[1];
        if (sink_j) sink_j[k] = out[2];
    }
    return rc;
}


/* ======================================================================================================================
 * Round 2: banded scoring and traceback (the rest of SURVEY row f4).  Restated from
 *   banded scoring     gotoh/gotoh_banded_inl.h:405-627 (gotoh_alignment_score_dispatch::run), sw/sw_banded_inl.h:341-512,
 *                      ed/ed_banded_inl.h (edit distance = SW with 0/-1/-1/-1); text cache: alignment_base_inl.h:75-99
 *   full traceback     alignment_inl.h:361-467 (driver), gotoh/gotoh_inl.h:513-555 (direction vectors), :1806-1872 (walker),
 *                      sw/sw_inl.h:476-500 (directions), :1650-1698 (walker)
 *   banded traceback   banded_inl.h:363-427 (driver), gotoh/gotoh_banded_inl.h:860-927, sw/sw_banded_inl.h:735-790 (walkers)
 * The reference's checkpoints / submatrices are a memory-saving device (it recomputes 32-row windows from short2 checkpoints);
 * inside the int16 parity domain they hold the same values as the full matrices kept here.
 * Pinned to the reference's own host build (oracle/ref_nvbio/nvbio_ref.cpp: nvbio_ref_banded_score / _traceback /
 * _banded_traceback) by tests/test_nvbio_oracle.py, including the known answers of the reference's own test program
 * (NvB/nvbio-test/alignment_test.cu:676-825).
 *
 * Banded geometry: row i = pattern symbol i (0-based), band entry j = 0..BAND-1 is text symbol i + j.  A move to the previous
 * row at the same text position ("top", band[j+1] of the previous row) consumes a pattern symbol; a move inside the row ("left")
 * consumes a text symbol.  NOTE (reference behaviour, reproduced): the banded Smith-Waterman prices the pattern-consuming move
 * with deletion() and the text-consuming one with insertion() -- the opposite of its full-matrix code (sw_banded_inl.h:378-379
 * vs sw_inl.h:484-485); the generic text cache (every BAND other than 3, 5, 7, 15) keeps 2 bits per symbol, so cached text symbols
 * are truncated to their low 2 bits (the 255 sentinel past the end of the text turns into 3 once it enters the cache).
 * Direction codes (alignment_base.h:139-150): 0 SUBSTITUTION, 1 INSERTION, 2 DELETION, 3 SINK, 4 INSERTION_EXT, 8 DELETION_EXT. */
enum { D_SUB = 0, D_INS = 1, D_DEL = 2, D_SINK = 3, D_IEXT = 4, D_DEXT = 8 };

typedef struct { int have; int best; uint32_t x, y; } best_sink;   /* BestSink (sink_inl.h:52-60): `score <= new` replaces */
static inline void sink_report(best_sink* s, int score, uint32_t x, uint32_t y) {
    if (!s->have || s->best <= score) { s->best = score; s->x = x; s->y = y; s->have = 1; }
}

static void scheme_costs(int algo, const int32_t* scheme, int* match, int* mismatch, int* go_t, int* ge_t, int* go_p, int* ge_p) {
    *match = scheme[0]; *mismatch = scheme[1];
    if (algo == 0) { *go_t = *go_p = scheme[2]; *ge_t = *ge_p = scheme[3]; }
    else if (algo == 1) { *go_t = *ge_t = scheme[2]; *go_p = *ge_p = scheme[3]; }     /* deletion, insertion */
    else { *match = 0; *mismatch = -1; *go_t = *ge_t = *go_p = *ge_p = -1; }
}

/* The banded DP.  dirs (optional): M * band direction bytes, dirs[i * band + j].  Returns 0 when the reference's function
 * returns false (text shorter than the pattern: nothing computed), 1 otherwise. */
static int banded_dp(int algo, int type, int band, const int32_t* scheme, const uint8_t* pattern, uint32_t M,
                     const uint8_t* text, uint32_t N, best_sink* sink, uint8_t* dirs) {
    int match, mismatch, go_t, ge_t, go_p, ge_p;
    scheme_costs(algo, scheme, &match, &mismatch, &go_t, &ge_t, &go_p, &ge_p);
    if (N < M) return 0;
    const int B = band;
    const int packed = !(B == 3 || B == 5 || B == 7 || B == 15);
    int H[64], F[64]; uint32_t cache[64];
    for (int j = 0; j < B - 1; ++j) cache[j] = packed ? (text[j] & 3u) : text[j];
    if (algo == 0) {
        const int Go = go_p, Ge = ge_p;
        const int mx = (Go > Ge ? Go : Ge);
        const int inf = -32768 - mx;
        H[0] = 0;
        for (int j = 1; j < B; ++j) H[j] = (type == 0) ? go_t + (j - 1) * ge_t : 0;
        for (int j = 0; j < B; ++j) F[j] = inf;
        for (uint32_t i = 0; i < M; ++i) {
            const uint8_t q = pattern[i];
            int hdir, edir = D_SUB, E;
            {   /* j == 0 */
                const int ftop = F[1] + Ge, htop = H[1] + Go;
                F[0] = max2(ftop, htop);
                const int fdir = ftop > htop ? D_DEXT : D_SUB;
                const int diag = H[0] + ((cache[0] == q) ? match : mismatch), top = F[0];
                int hi = max2(top, diag);
                hdir = D_SUB;
                if (type == 1) { hi = max2(hi, 0); hdir = (hi == 0) ? D_SINK : hdir; sink_report(sink, hi, i + 1, i + 1); }
                H[0] = hi;
                if (dirs) {
                    const int hd = (hdir == D_SINK) ? D_SINK : (top > diag ? D_INS : D_SUB);
                    dirs[(size_t)i * B + 0] = (uint8_t)(((type == 1 && hi == 0) ? D_SINK : hd) | D_SUB | fdir);
                }
            }
            E = H[0] + Go;
            for (int j = 1; j < B - 1; ++j) {
                const int ftop = F[j + 1] + Ge, htop = H[j + 1] + Go;
                F[j] = max2(ftop, htop);
                const int fdir = ftop > htop ? D_DEXT : D_SUB;
                const uint32_t g = cache[j]; cache[j - 1] = g;
                const int diag = H[j] + ((g == q) ? match : mismatch), top = F[j], left = E;
                int hi = max2(max2(top, left), diag);
                hdir = D_SUB;
                if (type == 1) { hi = max2(hi, 0); hdir = (hi == 0) ? D_SINK : hdir; sink_report(sink, hi, i + j + 1, i + 1); }
                H[j] = hi;
                if (dirs) {
                    const int hd = (hdir == D_SINK) ? D_SINK : (top > left ? (top > diag ? D_INS : D_SUB) : (left > diag ? D_DEL : D_SUB));
                    dirs[(size_t)i * B + j] = (uint8_t)(((type == 1 && hi == 0) ? D_SINK : hd) | edir | fdir);
                }
                const int eleft = E + Ge, ediag = hi + Go;
                edir = (eleft > ediag) ? D_IEXT : D_SUB;
                E = max2(ediag, eleft);
            }
            const uint32_t g = (i + B - 1 < N) ? text[i + B - 1] : 255u;
            cache[B - 2] = packed ? (g & 3u) : g;
            {   /* j == B - 1 */
                F[B - 1] = inf;
                const int diag = H[B - 1] + ((g == q) ? match : mismatch), left = E;
                int hi = max2(left, diag);
                hdir = D_SUB;
                if (type == 1) { hi = max2(hi, 0); hdir = (hi == 0) ? D_SINK : hdir; sink_report(sink, hi, i + B, i + 1); }
                H[B - 1] = hi;
                if (dirs) {
                    const int hd = (hdir == D_SINK) ? D_SINK : (left > diag ? D_DEL : D_SUB);
                    dirs[(size_t)i * B + B - 1] = (uint8_t)(((type == 1 && hi == 0) ? D_SINK : hd) | edir | D_SUB);
                }
            }
        }
    } else {
        const int G = go_t, I = go_p;                         /* deletion(), insertion() */
        for (int j = 0; j < B; ++j) H[j] = (type == 0) ? j * G : 0;
        for (uint32_t i = 0; i < M; ++i) {
            const uint8_t q = pattern[i];
            {
                const int diag = H[0] + ((cache[0] == q) ? match : mismatch), top = H[1] + G;
                int hi = max2(top, diag);
                if (type == 1) { hi = max2(hi, 0); sink_report(sink, hi, i + 1, i + 1); }
                H[0] = hi;
                if (dirs) dirs[(size_t)i * B + 0] = (uint8_t)(top > diag ? D_INS : D_SUB);
            }
            for (int j = 1; j < B - 1; ++j) {
                const uint32_t g = cache[j]; cache[j - 1] = g;
                const int diag = H[j] + ((g == q) ? match : mismatch), top = H[j + 1] + G, left = H[j - 1] + I;
                int hi = max2(max2(top, left), diag);
                if (type == 1) { hi = max2(hi, 0); sink_report(sink, hi, i + j + 1, i + 1); }
                H[j] = hi;
                if (dirs) dirs[(size_t)i * B + j] = (uint8_t)(top > left ? (top > diag ? D_INS : D_SUB) : (left > diag ? D_DEL : D_SUB));
            }
            const uint32_t g = (i + B - 1 < N) ? text[i + B - 1] : 255u;
            cache[B - 2] = packed ? (g & 3u) : g;
            {
                const int diag = H[B - 1] + ((g == q) ? match : mismatch), left = H[B - 2] + I;
                int hi = max2(left, diag);
                if (type == 1) { hi = max2(hi, 0); sink_report(sink, hi, i + B, i + 1); }
                H[B - 1] = hi;
                if (dirs) dirs[(size_t)i * B + B - 1] = (uint8_t)(left > diag ? D_DEL : D_SUB);
            }
        }
    }
    if (type == 0) sink_report(sink, H[B - 1], M + B - 1, M);
    else if (type == 2) {
        const uint32_t lim = (M + B - 1 < N ? M + B - 1 : N) - (M - 1);
        sink_report(sink, H[0], M, M);
        for (int j = 1; j < B; ++j) if ((uint32_t)j < lim) sink_report(sink, H[j], M + j, M);
    }
    return 1;
}

/* out[0] score, out[1] sink text coordinate, out[2] sink pattern coordinate (1-based ends; -1 when nothing was reported),
 * out[3] = the reference function's bool.  Needs 3 <= band <= 63, N >= band - 1 (the reference reads the first band - 1 text
 * symbols unguarded) and M >= 1. */
int nvbio_compat_banded_score(int algo, int type, int band, const int32_t* scheme, const uint8_t* pattern, uint32_t M,
                              const uint8_t* text, uint32_t N, int32_t* out) {
    if (band < 3 || band > 63 || N + 1 < (uint32_t)band) return -1;
    best_sink s = {0, -(1 << 30), (uint32_t)-1, (uint32_t)-1};   /* BestSink(): Field_traits<int32>::min() = -(1 << 30), numbers.h:832-836 */
    const int ok = banded_dp(algo, type, band, scheme, pattern, M, text, N, &s, NULL);
    out[0] = s.best; out[1] = (int32_t)s.x; out[2] = (int32_t)s.y; out[3] = ok;
    return 0;
}

/* ops: push order (from the END of the alignment to its source).  out[0] score, out[1..2] source (text, pattern),
 * out[3..4] sink (text, pattern), out[5] number of ops, out[6] clip at the end of the pattern, out[7] clip at its beginning,
 * out[8] number of clip() calls (2, or 0 when nothing was reported). */
static int push_op(uint8_t* ops, uint32_t cap, uint32_t* n, int op) { if (*n >= cap) return -1; ops[(*n)++] = (uint8_t)op; return 0; }

int nvbio_compat_banded_traceback(int algo, int type, int band, const int32_t* scheme, const uint8_t* pattern, uint32_t M,
                                  const uint8_t* text, uint32_t N, int32_t* out, uint8_t* ops, uint32_t cap) {
    if (band < 3 || band > 63 || N + 1 < (uint32_t)band) return -1;
    uint8_t* dirs = (uint8_t*)malloc((size_t)(M ? M : 1) * band);
    if (!dirs) return -1;
    best_sink s = {0, -(1 << 30), (uint32_t)-1, (uint32_t)-1};   /* BestSink(): Field_traits<int32>::min() = -(1 << 30), numbers.h:832-836 */
    banded_dp(algo, type, band, scheme, pattern, M, text, N, &s, dirs);
    uint32_t n = 0; int rc = 0;
    out[0] = s.best; out[3] = (int32_t)s.x; out[4] = (int32_t)s.y;
    if (s.x == (uint32_t)-1 || s.y == (uint32_t)-1) { out[1] = out[2] = out[3] = out[4] = -1; out[5] = out[6] = out[7] = out[8] = 0; free(dirs); return 0; }
    out[6] = (int32_t)(M - s.y);
    int entry = (int)s.x - (int)s.y, row = (int)s.y - 1, state = 0;       /* 0 H, 1 E, 2 F */
    int found = 0;
    while (row >= 0) {
        const int op = dirs[(size_t)row * band + entry], h_op = op & 3;
        if (algo == 0) {
            if (type == 1 && state == 0 && h_op == D_SINK) { found = 1; break; }
            if (state == 1) { if ((op & D_IEXT) == 0) state = 0; --entry; rc |= push_op(ops, cap, &n, D_DEL); }
            else if (state == 2) { if ((op & D_DEXT) == 0) state = 0; ++entry; --row; rc |= push_op(ops, cap, &n, D_INS); }
            else if (h_op == D_DEL) state = 1;
            else if (h_op == D_INS) state = 2;
            else { --row; rc |= push_op(ops, cap, &n, D_SUB); }
        } else {
            if (type == 1 && op == D_SINK) { found = 1; break; }
            if (op == D_DEL) { --entry; rc |= push_op(ops, cap, &n, D_DEL); }
            else if (op == D_INS) { ++entry; --row; rc |= push_op(ops, cap, &n, D_INS); }
            else { --row; rc |= push_op(ops, cap, &n, D_SUB); }
        }
        if (entry < 0 || entry >= band) { rc = -2; break; }                 /* the reference asserts this (undefined otherwise) */
    }
    uint32_t sy = found ? (uint32_t)(row + 1) : 0u, sx = (uint32_t)entry + sy;
    out[1] = (int32_t)sx; out[2] = (int32_t)sy; out[5] = (int32_t)n; out[7] = (int32_t)sy; out[8] = 2;
    free(dirs);
    return rc;
}

/* Full-matrix traceback.  `bandlen` is the reference's internal BAND_LEN (tie order of LOCAL sinks, as in nvbio_compat_score). */
int nvbio_compat_traceback(int algo, int type, const int32_t* scheme, const uint8_t* pattern, uint32_t M,
                           const uint8_t* text, uint32_t N, uint32_t bandlen, int32_t* out, uint8_t* ops, uint32_t cap) {
    int match, mismatch, go_t, ge_t, go_p, ge_p;
    scheme_costs(algo, scheme, &match, &mismatch, &go_t, &ge_t, &go_p, &ge_p);
    if (bandlen == 0) bandlen = 8;
    uint8_t* dirs = (uint8_t*)malloc((size_t)(N ? N : 1) * (M ? M : 1));
    int* H = (int*)malloc(sizeof(int) * (M + 1));
    int* F = (int*)malloc(sizeof(int) * (M + 1));
    if (!dirs || !H || !F) return -1;
    H[0] = 0;
    for (uint32_t j = 1; j <= M; ++j) { H[j] = (type == 1) ? 0 : go_p + ge_p * (int)(j - 1); F[j] = NEG; }
    int best = INT32_MIN; uint32_t bi = (uint32_t)-1, bj = (uint32_t)-1; uint64_t bkey = 0; int have = 0;
    for (uint32_t i = 1; i <= N; ++i) {
        int diag = H[0];
        H[0] = (type == 0) ? go_t + ge_t * (int)(i - 1) : 0;
        int E = (type == 1) ? 0 : NEG;
        for (uint32_t j = 1; j <= M; ++j) {
            int d;
            const int s = (text[i - 1] == pattern[j - 1]) ? match : mismatch;
            if (algo == 0) {
                const int ftop = F[j] + ge_t, htop = H[j] + go_t;
                F[j] = max2(ftop, htop);
                const int fdir = ftop > htop ? D_DEXT : D_SUB;
                const int eleft = E + ge_p, hleft = H[j - 1] + go_p;
                E = max2(eleft, hleft);
                const int edir = eleft > hleft ? D_IEXT : D_SUB;
                const int dg = diag + s, top = F[j], left = E;
                int hi = max2(max2(left, top), dg);
                if (type == 1) hi = max2(hi, 0);
                diag = H[j]; H[j] = hi;
                const int hd = top > left ? (top > dg ? D_DEL : D_SUB) : (left > dg ? D_INS : D_SUB);
                d = ((type == 1 && hi == 0) ? D_SINK : hd) | edir | fdir;
            } else {
                const int dg = diag + s, top = H[j] + go_t, left = H[j - 1] + go_p;
                int hi = max2(max2(top, left), dg);
                if (type == 1) hi = max2(hi, 0);
                diag = H[j]; H[j] = hi;
                const int hd = top > left ? (top > dg ? D_DEL : D_SUB) : (left > dg ? D_INS : D_SUB);
                d = (type == 1 && hi == 0) ? D_SINK : hd;
            }
            dirs[(size_t)(i - 1) * M + (j - 1)] = (uint8_t)d;
            if (type == 1) {
                const uint64_t key = ((uint64_t)((j - 1) / bandlen) << 40) | ((uint64_t)i << 20) | j;
                if (!have || H[j] > best || (H[j] == best && key > bkey)) { best = H[j]; bi = i; bj = j; bkey = key; have = 1; }
            }
        }
        if (type == 2 && (!have || H[M] >= best)) { best = H[M]; bi = i; bj = M; have = 1; }
    }
    if (type == 0 && N > 0) { best = H[M]; bi = N; bj = M; have = 1; }
    free(H); free(F);
    out[0] = best; out[3] = (int32_t)bi; out[4] = (int32_t)bj;
    if (bi == (uint32_t)-1 || bj == (uint32_t)-1) { out[1] = out[2] = -1; out[5] = out[6] = out[7] = out[8] = 0; free(dirs); return 0; }
    uint32_t n = 0; int rc = 0;
    out[6] = (int32_t)(M - bj);
    int row = (int)bi, col = (int)bj, state = 0;                             /* 1-based; the walk ends at row 0 or column 0 */
    while (row > 0 && col > 0) {
        const int op = dirs[(size_t)(row - 1) * M + (col - 1)], h_op = op & 3;
        if (algo == 0) {
            if (type == 1 && state == 0 && h_op == D_SINK) break;
            if (state == 1) { if ((op & D_IEXT) == 0) state = 0; --col; rc |= push_op(ops, cap, &n, D_INS); }
            else if (state == 2) { if ((op & D_DEXT) == 0) state = 0; --row; rc |= push_op(ops, cap, &n, D_DEL); }
            else if (h_op == D_INS) state = 1;
            else if (h_op == D_DEL) state = 2;
            else { --col; --row; rc |= push_op(ops, cap, &n, D_SUB); }
        } else {
            if (type == 1 && op == D_SINK) break;
            if (op != D_DEL) --col;
            if (op != D_INS) --row;
            rc |= push_op(ops, cap, &n, op);
        }
    }
    if ((type == 2 || type == 0) && row == 0) for (; col > 0; --col) rc |= push_op(ops, cap, &n, D_INS);
    if (type == 0 && col == 0) for (; row > 0; --row) rc |= push_op(ops, cap, &n, D_DEL);
    out[1] = row; out[2] = col; out[5] = (int32_t)n; out[7] = col; out[8] = 2;
    free(dirs);
    return rc;
}
