/*
 * gotoh_compat.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, scalar, int32 restatement of the GASAL2 batched pairwise-alignment path of
 * Genomics-GPU (reference: Non-CDP/GASAL2/src).  It exists only so that tests/, bench.py's
 * cpu_baseline leg and __graft_entry__.smoke() can check the CUDA engine; nothing in the product
 * path may call it.  Parity status: PINNED -- tests/test_oracle.py checks every function here
 * against (a) the reference's own kernel headers compiled for the host (oracle/_ref, built by
 * oracle/Makefile from /root/reference in place) and (b) the committed golden vectors under
 * tests/golden/ that were generated from that literal build (tests/golden/make_golden.py).
 *
 * Every quirk of the reference that is visible in its outputs is reproduced on purpose
 * (SURVEY.md section 8, Q1..Q14).  Citations are relative to Non-CDP/GASAL2/src/.
 *
 * Conventions (same as the reference): q = query index = column, t = target index = row,
 * sequences are 4-bit packed, 8 bases per uint32, first base in bits 31:28
 * (kernels/pack_rc_seqs.h:13-53); every sequence occupies whole words, padded with the N code.
 * The DP runs over the PADDED matrix in strips of 8 target rows (outer), query columns (inner).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GC_MINUS_INF (-32768)          /* gasal_kernels.h:35  MINUS_INF = SHRT_MIN */

enum { GC_WITHOUT_START = 0, GC_WITH_START = 1, GC_WITH_TB = 2 };              /* gasal.h:37-41 */
enum { GC_NONE = 0, GC_QUERY = 1, GC_TARGET = 2, GC_BOTH = 3 };                /* gasal.h:50-55 */
enum { GC_GLOBAL = 1, GC_SEMI_GLOBAL = 2, GC_LOCAL = 3, GC_BANDED = 5, GC_KSW = 6 }; /* gasal.h:57-65 */

typedef struct {
    int32_t match, mismatch, gap_open, gap_extend;   /* positive penalties, gasal.h:160-165 */
    int32_t n_code;                                   /* N_CODE, default 0x4E ('N') */
    int32_t n_penalty;                                /* <0: N_PENALTY undefined (gasal_kernels.h:39-56) */
} gc_scoring_t;

typedef struct { uint32_t x, y, z, w; } gc_uint4;

static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }

/* base k (0..7) of a packed word: first base in the top nibble */
static inline uint32_t nib(uint32_t w, int k) { return (w >> (28 - 4 * k)) & 15u; }
static inline uint32_t n_words(uint32_t len) { return (len >> 3) + ((len & 7) ? 1 : 0); }

/* gasal_kernels.h:39-56.  LOCAL flavour: any N scores 0 (or -N_PENALTY). */
static inline int32_t sub_local(const gc_scoring_t *sc, uint32_t r, uint32_t g)
{
    int32_t s = (r == g) ? sc->match : -sc->mismatch;
    uint32_t nv = (uint32_t)sc->n_code & 15u;
    if (r == nv || g == nv) s = (sc->n_penalty >= 0) ? -sc->n_penalty : 0;
    return s;
}
/* GLOBAL flavour: N is an ordinary symbol unless N_PENALTY is defined. */
static inline int32_t sub_global(const gc_scoring_t *sc, uint32_t r, uint32_t g)
{
    int32_t s = (r == g) ? sc->match : -sc->mismatch;
    if (sc->n_penalty >= 0) {
        uint32_t nv = (uint32_t)sc->n_code & 15u;
        if (r == nv || g == nv) s = -sc->n_penalty;
    }
    return s;
}

/* ------------------------------------------------------------------------------------------ */
/* 4-bit packing: kernels/pack_rc_seqs.h:13-53 (nibble = byte & 15, first byte -> top nibble)  */
void gc_pack(const uint8_t *ascii, uint32_t n_bytes, uint32_t *words)
{
    for (uint32_t w = 0; w < n_bytes / 8; w++) {
        uint32_t v = 0;
        for (int k = 0; k < 8; k++) v |= (uint32_t)(ascii[8 * w + k] & 15u) << (28 - 4 * k);
        words[w] = v;
    }
}

static inline uint32_t rev_nibbles(uint32_t w)
{
    uint32_t r = 0;
    for (int k = 0; k < 8; k++) r |= ((w >> (4 * k)) & 15u) << (28 - 4 * k);
    return r;
}

/* kernels/pack_rc_seqs.h:56-212.  op bit0 = reverse, bit1 = complement (A<->T = 1<->4, C<->G = 3<->7).
 * The reverse branch has UNDEFINED BEHAVIOUR in C: the pad counter compares a nibble with the full
 * N_CODE 0x4E, so nbr_N is always 0 and the code shifts 32-bit words by 32 (:136,153).  Two readings:
 *   mode 0 (GC_REVCOMP_PADDED): shifts by 32 give 0 (PTX shl/shr semantics) -> a clean nibble reversal
 *          of the whole PADDED word array, pads end up in front.  Engine default.
 *   mode 1 (GC_REVCOMP_REF_NVCC129): what nvcc 12.9 actually emits for sm_100 (it folds the undefined
 *          operand to 0; SASS in profiles/reference_revcomp_kernel_sm100_nvcc12.9.sass): for
 *          i < ceil(nw/2): out[nw-1-i] = reverse(in[i]), out[i] = 0 (the middle word keeps its reversal);
 *          the second half of the input is lost.  Verified against the reference binary on a B200
 *          (tests/test_ref_gpu.py). */
static int g_revcomp_mode = 0;
void gc_set_revcomp_mode(int mode) { g_revcomp_mode = mode; }

void gc_revcomp_one(uint32_t *words, uint32_t len, uint8_t op)
{
    uint32_t nw = n_words(len);
    if (op & 1) {
        if (g_revcomp_mode == 1) {
            uint32_t swap = (nw >> 1) + (nw & 1);
            for (uint32_t i = 0; i < swap; i++) {
                uint32_t a = rev_nibbles(words[i]);
                words[i] = 0;
                words[nw - 1 - i] = a;
            }
        } else {
            for (uint32_t i = 0; i < nw / 2; i++) {
                uint32_t a = rev_nibbles(words[i]), b = rev_nibbles(words[nw - 1 - i]);
                words[i] = b; words[nw - 1 - i] = a;
            }
            if (nw & 1) words[nw / 2] = rev_nibbles(words[nw / 2]);
        }
    }
    if (op & 2) {
        for (uint32_t i = 0; i < nw; i++) {
            uint32_t w = words[i], o = 0;
            for (int k = 0; k < 8; k++) {
                uint32_t c = nib(w, k);
                c = (c == 1) ? 4 : (c == 4) ? 1 : (c == 3) ? 7 : (c == 7) ? 3 : c;
                o |= c << (28 - 4 * k);
            }
            words[i] = o;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Column state carried between strips: (H of the strip's last row, E entering the next strip), */
/* parked as int16 exactly like the reference's `short2 global[]` (Q7: silent truncation).      */
typedef struct { int16_t h, e; } colstate_t;

typedef struct {
    int32_t score, q_end, t_end, q_start, t_start;
    int32_t score2, q_end2, t_end2;
} gc_result_t;

/* Direction nibble of one cell: kernels/local_kernel_template.h:45-60, global.h:14-26.
 * bits1:0 = source of H (0 diag/match, 1 diag/mismatch, 3 from F, 2 otherwise),
 * bit2 = E extended (not opened), bit3 = F extended. */
static inline uint32_t tb_nibble(int32_t d, int32_t diag, int32_t h, int32_t f, int32_t e, int32_t oe, int32_t ext)
{
    uint32_t m_or_x = (d >= diag) ? 0u : 1u;
    uint32_t code = (h == d) ? m_or_x : ((h == f) ? 3u : 2u);
    uint32_t fbit = ((d - oe) > (f - ext)) ? 0u : 1u;
    uint32_t ebit = ((d - oe) > (e - ext)) ? 0u : 1u;
    return code | (ebit << 2) | (fbit << 3);
}

static inline void tb_store(gc_uint4 *tb, uint64_t tile_no, uint64_t n_tasks, uint64_t tid, int col_in_tile, uint32_t word)
{
    gc_uint4 *t = &tb[tile_no * n_tasks + tid];
    if (col_in_tile == 0) t->x = word; else if (col_in_tile == 1) t->y = word; else if (col_in_tile == 2) t->z = word; else t->w = word;
}

/* ------------------------------------------------------------------------------------------ */
/* LOCAL: kernels/local_kernel_template.h:71-519.  D-based gap opening, clamp at 0, all        */
/* boundaries 0; first cell with the maximum in (strip, column, row-in-strip) order wins (Q2). */
static void local_one(const gc_scoring_t *sc, const uint32_t *qw, const uint32_t *tw, uint32_t qlen, uint32_t tlen,
                      int start_mode, int second_best, colstate_t *cs,
                      gc_uint4 *tb, uint64_t n_tasks, uint64_t tid, gc_result_t *out)
{
    const int32_t oe = sc->gap_open + sc->gap_extend, ext = sc->gap_extend;
    const uint32_t QR = n_words(qlen), TR = n_words(tlen);
    int32_t best = 0, best_t = 0, prev_best = 0, best_q = 0;
    int32_t best2 = 0, prev_best2 = 0, best2_q = 0, best2_t = 0;
    uint64_t tile_no = 0;

    for (uint32_t c = 0; c < QR * 8; c++) { cs[c].h = 0; cs[c].e = 0; }        /* :118-120 */

    for (uint32_t i = 0; i < TR; i++) {                                           /* :122 */
        int32_t diag[8], F[8];
        for (int r = 0; r < 8; r++) { diag[r] = 0; F[r] = 0; }                   /* :123-127 */
        const uint32_t gpac = tw[i];
        for (uint32_t c = 0; c < QR * 8; c++) {
            const uint32_t rb = nib(qw[c >> 3], c & 7);
            int32_t up = cs[c].h, e = cs[c].e;                                    /* :390-392 */
            uint32_t dirword = 0;
            for (int r = 0; r < 8; r++) {                                         /* CORE_LOCAL_COMPUTE :19-30 */
                const int32_t s = sub_local(sc, rb, nib(gpac, r));
                const int32_t d = diag[r] + s;
                int32_t h = imax(imax(imax(d, F[r]), e), 0);
                if (start_mode == GC_WITH_TB) dirword |= tb_nibble(d, diag[r], h, F[r], e, oe, ext) << (28 - 4 * r);
                F[r] = imax(d - oe, F[r] - ext);
                e = imax(d - oe, e - ext);
                if (best < h) { best_t = (int32_t)(8 * i + r); best = h; }
                if (second_best && best2 < h && best > h) { best2_t = (int32_t)(8 * i + r); best2 = h; }  /* :147-149 */
                diag[r] = up;                                                     /* p[m] = h[m-1] */
                up = h;
            }
            cs[c].h = (int16_t)up; cs[c].e = (int16_t)e;                          /* :406-408 */
            if (start_mode == GC_WITH_TB) {
                tb_store(tb, tile_no, n_tasks, tid, (int)(c & 3), dirword);       /* :257,380 */
                if ((c & 3) == 3) tile_no++;
            }
            if (prev_best < best) best_q = (int32_t)c;                            /* :412 */
            if (second_best) {
                if (prev_best2 < best) best2_q = (int32_t)c;                      /* :162 (sic: compares with maxHH) */
                prev_best2 = imax(best2, prev_best2);
            }
            prev_best = imax(best, prev_best);
        }
    }
    out->score = best; out->q_end = best_q; out->t_end = best_t;
    out->score2 = best2; out->q_end2 = best2_q; out->t_end2 = best2_t;

    if (start_mode != GC_WITH_START) return;

    /* reverse pass, :441-511: same DP backwards from the END OF THE WORDS holding the end cell,
     * stop as soon as max >= fwd score, checked only between columns. */
    const int32_t fwd = best;
    const int32_t rend_reg = imin((best_q >> 3) + 1, (int32_t)QR);
    const int32_t gend_reg = imin((best_t >> 3) + 1, (int32_t)TR);
    int32_t m = 0, pm = 0, sx = 0, sy = 0;
    for (uint32_t c = 0; c < QR * 8; c++) { cs[c].h = 0; cs[c].e = 0; }
    int32_t gidx = ((gend_reg << 3) + 8) - 1;
    for (int32_t i = 0; i < gend_reg && m < fwd; i++) {
        int32_t diag[8], F[8];
        for (int r = 0; r < 8; r++) { diag[r] = 0; F[r] = 0; }
        const uint32_t gpac = tw[gend_reg - 1 - i];
        gidx -= 8;
        int32_t ridx = (rend_reg << 3) - 1;
        int32_t gcol = 0;
        for (int32_t j = 0; j < rend_reg && m < fwd; j++) {
            const uint32_t rpac = qw[rend_reg - 1 - j];
            for (int k = 7; k >= 0 && m < fwd; k--) {                            /* lowest nibble first */
                const uint32_t rb = nib(rpac, k);
                int32_t up = cs[gcol].h, e = cs[gcol].e;
                for (int r = 0; r < 8; r++) {                                     /* row r reads target nibble 7-r */
                    const int32_t s = sub_local(sc, rb, nib(gpac, 7 - r));
                    const int32_t d = diag[r] + s;
                    int32_t h = imax(imax(imax(d, F[r]), e), 0);
                    F[r] = imax(d - oe, F[r] - ext);
                    e = imax(d - oe, e - ext);
                    if (m < h) { sy = gidx + r; m = h; }                         /* Q3: gidx + (m-1), mirrored */
                    diag[r] = up; up = h;
                }
                cs[gcol].h = (int16_t)up; cs[gcol].e = (int16_t)e;
                if (pm < m) sx = ridx;
                pm = imax(m, pm);
                ridx--; gcol++;
            }
        }
    }
    out->q_start = sx; out->t_start = sy;
}

/* ------------------------------------------------------------------------------------------ */
/* SEMI-GLOBAL: kernels/semiglobal_kernel_template.h:39-388.  H-based opening, no clamp.       */
typedef struct { int32_t best, x, y, best2, x2, y2; } sg_track_t;

/* one full DP pass (forward, or the reverse pass over reversed sequences). first_strip/stop are
 * for the WITH_START pass (:273,301,323). Column state array `cs` has `cs_len` entries. */
static void semiglobal_pass(const gc_scoring_t *sc, const uint32_t *qw, const uint32_t *tw, uint32_t qlen, uint32_t tlen,
                            int head, int tail, int second_best, colstate_t *cs, uint32_t cs_len,
                            uint32_t first_strip, int use_stop, int32_t stop_score, sg_track_t *tr)
{
    const int32_t o = sc->gap_open, ext = sc->gap_extend, oe = o + ext;
    const uint32_t QR = n_words(qlen), TR = n_words(tlen);
    const int head_q = (head == GC_QUERY || head == GC_BOTH);      /* top row free    :87 */
    const int head_t = (head == GC_TARGET || head == GC_BOTH);     /* left column free :114 */

    if (head_q) { for (uint32_t c = 0; c < cs_len; c++) { cs[c].h = 0; cs[c].e = 0; } }
    else {
        cs[0].h = 0; cs[0].e = (int16_t)GC_MINUS_INF;
        for (uint32_t c = 1; c < cs_len; c++) { cs[c].h = (int16_t)(-(o + ext * (int32_t)c)); cs[c].e = (int16_t)GC_MINUS_INF; }
    }
    int32_t row_ctr = 0;           /* the reference's u-1 == r-1 counters, restarted per pass (:101-107) */

    for (uint32_t i = first_strip; i < TR && (!use_stop || tr->best < stop_score); i++) {
        int32_t Hleft[8], F[8], diag[8];
        if (head_t) { for (int r = 0; r < 8; r++) { Hleft[r] = 0; F[r] = GC_MINUS_INF; diag[r] = 0; } }
        else {
            for (int r = 0; r < 8; r++, row_ctr++) {
                Hleft[r] = -(o + ext * row_ctr);                                       /* :125 */
                F[r] = GC_MINUS_INF;
                diag[r] = (row_ctr == 0) ? 0 : -(o + ext * row_ctr);                   /* :127 */
            }
        }
        const uint32_t gpac = tw[i];
        uint32_t c = 0;
        for (uint32_t j = 0; j < QR && (!use_stop || tr->best < stop_score); j++) {
            const uint32_t rpac = qw[j];
            for (int k = 0; k < 8; k++, c++) {
                const uint32_t rb = nib(rpac, k);
                int32_t up = cs[c].h, e = cs[c].e;
                int32_t hcol[8];
                for (int r = 0; r < 8; r++) {                                          /* CORE_COMPUTE_SEMIGLOBAL :17-28 */
                    const int32_t s = sub_local(sc, rb, nib(gpac, r));
                    F[r] = imax(Hleft[r] - oe, F[r] - ext);
                    int32_t h = diag[r] + s;
                    h = imax(h, F[r]);
                    e = imax(up - oe, e - ext);
                    h = imax(h, e);
                    Hleft[r] = h; hcol[r] = h;
                    diag[r] = up;
                    up = h;
                }
                cs[c].h = (int16_t)up; cs[c].e = (int16_t)e;
                if ((tail == GC_TARGET || tail == GC_BOTH) && c + 1 == qlen) {         /* :160-178 */
                    for (int r = 0; r < 8; r++) {
                        const int32_t t = (int32_t)(8 * i + r);
                        const int in = (uint32_t)t < tlen;
                        if (hcol[r] > tr->best && in) { tr->y = t; tr->best = hcol[r]; }
                        if (second_best && hcol[r] > tr->best2 && hcol[r] < tr->best && in) { tr->y2 = t; tr->best2 = hcol[r]; }
                    }
                }
            }
        }
    }
    if (tail == GC_QUERY || tail == GC_BOTH) {                                          /* :185-216 */
        for (uint32_t m = 0; m < cs_len; m++) {
            const int32_t v = cs[m].h;
            if (v > tr->best && m < qlen) { tr->x = (int32_t)m; tr->best = v; }
            if (second_best && v > tr->best2 && v < tr->best && m < tlen) { tr->x2 = (int32_t)m; tr->best2 = v; }   /* (sic) m < ref_len */
        }
        if (tr->x != (int32_t)tlen) tr->y = (int32_t)qlen;
        if (second_best && tr->x2 != (int32_t)tlen) tr->y2 = (int32_t)qlen;
    }
}

static void reverse_true_len(const uint32_t *w, uint32_t len, uint32_t *out, uint32_t out_words)
{
    memset(out, 0, sizeof(uint32_t) * out_words);                                       /* :237-242 zero pads */
    for (uint32_t k = 0; k < len; k++) {                                                /* :245-266 */
        const uint32_t i = len - 1 - k;
        out[k >> 3] |= nib(w[i >> 3], (int)(i & 7)) << (28 - 4 * (k & 7));
    }
}

/* `max_query_len` mirrors the reference's compile-time MAX_QUERY_LEN: it bounds the TAIL=QUERY/BOTH
 * scan (:187) and sizes the column-state array. Pass >= max(Q8, tlen) (Q10). */
static void semiglobal_one(const gc_scoring_t *sc, const uint32_t *qw, const uint32_t *tw, uint32_t qlen, uint32_t tlen,
                           int start_mode, int second_best, int head, int tail, uint32_t max_query_len,
                           colstate_t *cs, uint32_t *rq, uint32_t *rt, gc_result_t *out)
{
    sg_track_t tr;
    tr.best = GC_MINUS_INF; tr.x = (int32_t)tlen; tr.y = (int32_t)qlen;                  /* :63-64 (names swapped) */
    tr.best2 = GC_MINUS_INF; tr.x2 = (int32_t)tlen; tr.y2 = (int32_t)qlen;
    semiglobal_pass(sc, qw, tw, qlen, tlen, head, tail, second_best, cs, max_query_len, 0, 0, 0, &tr);
    out->score = tr.best; out->t_end = tr.y; out->q_end = tr.x;
    out->score2 = tr.best2; out->t_end2 = tr.y2; out->q_end2 = tr.x2;
    if (start_mode != GC_WITH_START) return;

    const uint32_t TR = n_words(tlen);
    reverse_true_len(qw, qlen, rq, max_query_len >> 3);
    reverse_true_len(tw, tlen, rt, max_query_len >> 3);
    const int32_t fwd = tr.best;
    /* :273 -- target_batch_regs is unsigned in the reference, so the subtraction wraps and the
     * strip loop (`i < target_batch_regs`, unsigned compare) simply does not run when it does. */
    const uint32_t ua = TR - (uint32_t)((tr.y >> 3) + 1);
    const uint32_t gend_reg = ua > 0 ? ua - 1 : ua;
    tr.best = GC_MINUS_INF; tr.y = 0;                                                    /* :275-276, x keeps fwd value */
    semiglobal_pass(sc, rq, rt, qlen, tlen, head, tail, 0, cs, max_query_len, gend_reg, 1, fwd, &tr);
    out->t_start = ((int32_t)tlen - 1) - tr.y;                                           /* :380-381 */
    out->q_start = ((int32_t)qlen - 1) - tr.x;
}

/* ------------------------------------------------------------------------------------------ */
/* GLOBAL: kernels/global.h:30-303.  D-based opening, no clamp, N is an ordinary symbol (Q4).  */
static void global_one(const gc_scoring_t *sc, const uint32_t *qw, const uint32_t *tw, uint32_t qlen, uint32_t tlen,
                       int start_mode, colstate_t *cs, gc_uint4 *tb, uint64_t n_tasks, uint64_t tid, gc_result_t *out)
{
    const int32_t o = sc->gap_open, ext = sc->gap_extend, oe = o + ext;
    const uint32_t QR = n_words(qlen), TR = n_words(tlen);
    int32_t last_col[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t tile_no = 0;
    cs[0].h = 0; cs[0].e = (int16_t)GC_MINUS_INF;                                        /* :57-60 */
    for (uint32_t c = 1; c < QR * 8; c++) { cs[c].h = (int16_t)(-(o + ext * (int32_t)c)); cs[c].e = (int16_t)GC_MINUS_INF; }
    int32_t row = 0;
    for (uint32_t i = 0; i < TR; i++) {
        int32_t diag[8], F[8];
        for (int r = 0; r < 8; r++, row++) { F[r] = GC_MINUS_INF; diag[r] = (row == 0) ? 0 : -(o + ext * row); }   /* :66-70 */
        const uint32_t gpac = tw[i];
        for (uint32_t c = 0; c < QR * 8; c++) {
            const uint32_t rb = nib(qw[c >> 3], c & 7);
            int32_t up = cs[c].h, e = cs[c].e;
            int32_t hcol[8];
            uint32_t dirword = 0;
            for (int r = 0; r < 8; r++) {                                                /* CORE_GLOBAL_COMPUTE :4-12 */
                const int32_t s = sub_global(sc, rb, nib(gpac, r));
                const int32_t d = diag[r] + s;
                int32_t h = imax(imax(d, F[r]), e);
                if (start_mode == GC_WITH_TB) dirword |= tb_nibble(d, diag[r], h, F[r], e, oe, ext) << (28 - 4 * r);
                F[r] = imax(d - oe, F[r] - ext);
                e = imax(d - oe, e - ext);
                hcol[r] = h; diag[r] = up; up = h;
            }
            cs[c].h = (int16_t)up; cs[c].e = (int16_t)e;
            if (start_mode == GC_WITH_TB) {
                tb_store(tb, tile_no, n_tasks, tid, (int)(c & 3), dirword);
                if ((c & 3) == 3) tile_no++;
            }
            if (c + 1 == qlen) for (int r = 0; r < 8; r++) last_col[r] = hcol[r];      /* :285-290 */
        }
    }
    /* :299  max_h[8 - (T8 - ref_len)], 1-based rows -> 0-based index 7 - (T8 - tlen) */
    out->score = last_col[7 - (int)(TR * 8 - tlen)];
}

/* ------------------------------------------------------------------------------------------ */
/* BANDED: kernels/banded.h:10-139.  H-based opening with clamp; band in whole 8x8 tiles.      */
static void banded_one(const gc_scoring_t *sc, const uint32_t *qw, const uint32_t *tw, uint32_t qlen, uint32_t tlen,
                       int32_t k_band, colstate_t *cs, gc_result_t *out)
{
    const int32_t ext = sc->gap_extend, oe = sc->gap_open + ext;
    const int32_t QR = (int32_t)n_words(qlen), TR = (int32_t)n_words(tlen);
    const int32_t kw = k_band >> 3;                                                      /* gasal_align.h:80 */
    const int32_t k_other = TR - (QR - kw);                                              /* :35 */
    int32_t best = 0, prev_best = 0, bx = 0, by = 0;
    for (int32_t c = 0; c < QR * 8; c++) { cs[c].h = 0; cs[c].e = 0; }
    for (int32_t i = 0; i < TR; i++) {
        int32_t Hleft[8], F[8], diag[8];
        for (int r = 0; r < 8; r++) { Hleft[r] = 0; F[r] = 0; diag[r] = 0; }
        const uint32_t gpac = tw[i];
        int32_t c = imax(0, i - k_other + 1) << 3;                                       /* :83 */
        const int32_t last_tile = imin(kw + i, QR);                                      /* :84 */
        for (int32_t j = c >> 3; j < last_tile; j++) {
            const uint32_t rpac = qw[j];
            for (int k = 0; k < 8; k++, c++) {
                const uint32_t rb = nib(rpac, k);
                int32_t up = cs[c].h, e = cs[c].e;
                for (int r = 0; r < 8; r++) {                                            /* :99-114 */
                    const int32_t s = sub_local(sc, rb, nib(gpac, r));
                    F[r] = imax(Hleft[r] - oe, F[r] - ext);
                    int32_t h = diag[r] + s;
                    h = imax(h, F[r]);
                    h = imax(h, 0);
                    e = imax(up - oe, e - ext);
                    h = imax(h, e);
                    if (best < h) { by = 8 * i + r; best = h; }
                    Hleft[r] = h; diag[r] = up; up = h;
                }
                cs[c].h = (int16_t)up; cs[c].e = (int16_t)e;
                if (prev_best < best) bx = c;
                prev_best = imax(best, prev_best);
            }
        }
    }
    out->score = best; out->q_end = bx; out->t_end = by;
}

/* ------------------------------------------------------------------------------------------ */
/* BANDED + WITH_TB (an EXTENSION: the reference has no banded traceback).  */
/* Definition: the LOCAL recurrence and direction nibbles (local_kernel_template.h:19-60) computed only */
/* inside the tile band of banded.h:35,83-85; cells outside the band read as 0, exactly as in the      */
/* banded kernel.  Wherever the band does not cut the optimal alignment this equals the reference's     */
/* LOCAL WITH_TB result (tests pin it that way).  The oracle keeps the full-matrix tile layout so the   */
/* walker below is shared; the engine stores only the band.                                             */
static void banded_tb_one(const gc_scoring_t *sc, const uint32_t *qw, const uint32_t *tw, uint32_t qlen, uint32_t tlen,
                          int32_t k_band, colstate_t *cs, gc_uint4 *tb, uint64_t n_tasks, uint64_t tid, gc_result_t *out)
{
    const int32_t ext = sc->gap_extend, oe = sc->gap_open + ext;
    const int32_t QR = (int32_t)n_words(qlen), TR = (int32_t)n_words(tlen);
    const int32_t kw = k_band >> 3;
    const int32_t k_other = TR - (QR - kw);
    int32_t best = 0, prev_best = 0, bx = 0, by = 0;
    for (int32_t c = 0; c < QR * 8; c++) { cs[c].h = 0; cs[c].e = 0; }
    for (int32_t i = 0; i < TR; i++) {
        int32_t F[8], diag[8];
        for (int r = 0; r < 8; r++) { F[r] = 0; diag[r] = 0; }
        const uint32_t gpac = tw[i];
        int32_t c = imax(0, i - k_other + 1) << 3;
        const int32_t last_tile = imin(kw + i, QR);
        for (int32_t j = c >> 3; j < last_tile; j++) {
            const uint32_t rpac = qw[j];
            for (int k = 0; k < 8; k++, c++) {
                const uint32_t rb = nib(rpac, k);
                int32_t up = cs[c].h, e = cs[c].e;
                uint32_t dirword = 0;
                for (int r = 0; r < 8; r++) {
                    const int32_t s = sub_local(sc, rb, nib(gpac, r));
                    const int32_t d = diag[r] + s;
                    int32_t h = imax(imax(imax(d, F[r]), e), 0);
                    dirword |= tb_nibble(d, diag[r], h, F[r], e, oe, ext) << (28 - 4 * r);
                    F[r] = imax(d - oe, F[r] - ext);
                    e = imax(d - oe, e - ext);
                    if (best < h) { by = 8 * i + r; best = h; }
                    diag[r] = up; up = h;
                }
                cs[c].h = (int16_t)up; cs[c].e = (int16_t)e;
                tb_store(tb, ((uint64_t)i * QR + j) * 2 + (k >> 2), n_tasks, tid, k & 3, dirword);
                if (prev_best < best) bx = c;
                prev_best = imax(best, prev_best);
            }
        }
    }
    out->score = best; out->q_end = bx; out->t_end = by;
}

/* ------------------------------------------------------------------------------------------ */
/* Traceback walker: kernels/get_tb.h:4-149.  Emits (count<<2)|op bytes in REVERSE alignment   */
/* order; op 0=M 1=X 2=D 3=I; runs capped at 63.  Returns n_ops; LOCAL also yields the starts. */
static uint32_t get_tb_one(const gc_scoring_t *sc, int is_local, const gc_uint4 *tb, uint64_t n_tasks, uint64_t tid,
                           uint32_t qlen, uint32_t tlen, gc_result_t *res, uint8_t *cigar)
{
    int i, j, total = 0, curr = 0;
    if (is_local) { i = res->t_end; j = res->q_end; total = res->score; }
    else { i = (int)tlen; j = (int)qlen; }                                               /* :21-24 (Q5) */
    const int q8 = (int)(n_words(qlen) * 8);
    const int oe = sc->gap_open + sc->gap_extend;
    uint32_t prev_op = 0, op_fill = 0;
    int n_ops = 0, count = 0, op_select = 3, op_shift = 0;
    uint32_t off = 0;
    while (i >= 0 && j >= 0) {
        const int cell = (((i >> 3) * q8) << 3) + (j << 3) + (i & 7);                    /* :50 */
        const int tile_no = cell >> 5;
        const gc_uint4 tile = tb[(uint64_t)tile_no * n_tasks + tid];
        const int in_tile = cell - (tile_no << 5);
        const int reg_no = in_tile >> 3, cell_in_reg = in_tile - (reg_no << 3);
        const uint32_t reg = reg_no == 0 ? tile.x : (reg_no == 1 ? tile.y : (reg_no == 2 ? tile.z : tile.w));
        const uint32_t cell_op = (reg >> (28 - (cell_in_reg << 2))) & 15u;
        const uint32_t op = (cell_op >> op_shift) & (uint32_t)op_select;
        op_fill = (op == 0 || op_select == 3) ? op : (uint32_t)op_shift;                 /* :74-82 */
        op_select = (op == 0 || (op == 1 && op_select == 3)) ? 3 : 1;
        op_shift = (op == 0 || (op == 1 && op_select == 3)) ? 0 : ((op == 2 || op == 3) ? (int)op : op_shift);
        if (count < 63 && op_fill == prev_op) count++;
        else {
            if (count > 0) { cigar[off++] = (uint8_t)(prev_op | (uint32_t)(count << 2)); n_ops++; }
            count = 1;
        }
        if (is_local) {                                                                   /* :100-103 */
            curr += ((op_fill == 2 || op_fill == 3) && prev_op != op_fill) ? -oe
                  : ((op_fill == 2 || op_fill == 3) ? -sc->gap_extend : (op_fill == 1 ? -sc->mismatch : sc->match));
            if (curr == total) break;
        }
        prev_op = op_fill;
        if (op_fill == 0 || op_fill == 1 || op_fill == 2) i--;
        if (op_fill == 0 || op_fill == 1 || op_fill == 3) j--;
    }
    /* note: on `break` prev_op still holds the previous op (as in the reference, :117) */
    cigar[off++] = (uint8_t)(prev_op | (uint32_t)(count << 2)); n_ops++;
    if (!is_local) {                                                                      /* :119-139 */
        while (i >= 0) { int c = (i + 1) <= 63 ? (i + 1) : 63; cigar[off++] = (uint8_t)(2 | (c << 2)); n_ops++; i -= 63; }
        while (j >= 0) { int c = (j + 1) <= 63 ? (j + 1) : 63; cigar[off++] = (uint8_t)(3 | (c << 2)); n_ops++; j -= 63; }
    } else { res->t_start = i; res->q_start = j; }                                        /* :142-145 */
    return (uint32_t)n_ops;
}

/* ------------------------------------------------------------------------------------------ */
/* KSW (seed extension): kernels/ksw_kernel_template.h:46-199 (BWA ksw_extend, z-drop off).    */
static void ksw_one(const gc_scoring_t *sc, const uint32_t *qw, const uint32_t *tw, uint32_t qlen_u, uint32_t tlen_u,
                    uint32_t h0u, int32_t *eh_h, int32_t *eh_e, uint32_t eh_len, gc_result_t *out)
{
    const int qlen = (int)qlen_u, tlen = (int)tlen_u, h0 = (int)h0u;
    const int e_del = sc->gap_extend, e_ins = sc->gap_extend, oe_del = sc->gap_open + e_del, oe_ins = sc->gap_open + e_ins;
    const uint32_t QRk = (qlen_u >> 3) + 1, TRk = (tlen_u >> 3) + 1;                     /* :57-58 (always +1) */
    for (uint32_t i = 0; i < eh_len; i++) { eh_h[i] = 0; eh_e[i] = 0; }
    eh_h[0] = h0; eh_h[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    int i, j;
    for (j = 2; j <= qlen && eh_h[j - 1] > e_ins; ++j) eh_h[j] = eh_h[j - 1] - e_ins;
    int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, beg = 0, end = qlen;
    for (uint32_t tt = 0; tt < TRk; tt++) {
        const uint32_t gpac = tw[tt];
        for (int tb = 0; tb < 8; tb++) {
            i = (int)tt * 8 + tb;
            if (i >= tlen) break;
            const uint32_t gbase = nib(gpac, tb);
            int f = 0, h1, m = 0, mj = -1;
            if (beg == 0) { h1 = h0 - (sc->gap_open + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
            j = 0;
            for (uint32_t qt = 0; qt < QRk; qt++) {
                const uint32_t rpac = qw[qt];
                for (int qb = 0; qb < 8; qb++) {
                    j = (int)qt * 8 + qb;
                    if (j < beg) continue;
                    if (j >= end) break;                                                  /* leaves the 8-base loop only (:117) */
                    const uint32_t rbase = nib(rpac, qb);
                    int M = eh_h[j], e = eh_e[j];
                    eh_h[j] = h1;
                    M = M ? M + sub_local(sc, rbase, gbase) : 0;
                    int h = M > e ? M : e; h = h > f ? h : f;
                    h1 = h;
                    mj = m > h ? mj : j;
                    m = m > h ? m : h;
                    int t = M - oe_del; t = t > 0 ? t : 0; e -= e_del; e = e > t ? e : t; eh_e[j] = e;
                    t = M - oe_ins; t = t > 0 ? t : 0; f -= e_ins; f = f > t ? f : t;
                }
            }
            eh_h[end] = h1; eh_e[end] = 0;
            if (j == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
            if (m == 0) break;                                                            /* :150 leaves the 8-row loop only */
            if (m > max) { max = m; max_i = i; max_j = mj; }
            for (j = beg; j < end && eh_h[j] == 0 && eh_e[j] == 0; ++j) ;
            beg = j;
            for (j = end; j >= beg && eh_h[j] == 0 && eh_e[j] == 0; --j) ;
            end = j + 2 < qlen ? j + 2 : qlen;
        }
    }
    if (gscore <= 0 || gscore <= max - 5) { out->score = max; out->q_end = max_j + 1; out->t_end = max_i + 1; }
    else { out->score = gscore; out->q_end = qlen; out->t_end = max_ie + 1; }
}

/* ------------------------------------------------------------------------------------------ */
/* Batch driver: the CPU equivalent of gasal_kernel_launcher (gasal_align.cu:10-25,            */
/* gasal_align.h:7-107): one "thread" per pair, outputs SoA arrays like gasal_res_t.           */
typedef struct {
    int32_t algo, start_mode, second_best, head, tail, k_band;
    uint32_t max_query_len;       /* 0 = derive from the batch; mirrors compile-time MAX_QUERY_LEN */
} gc_params_t;

typedef struct {
    int32_t *aln_score, *query_batch_end, *target_batch_end, *query_batch_start, *target_batch_start;
    int32_t *aln_score2, *query_batch_end2, *target_batch_end2;
    uint8_t *cigar;               /* WITH_TB: Q-batch-bytes long, written at query_offsets[i] (Q6) */
    uint32_t *n_cigar_ops;
} gc_results_t;

int gc_align_batch(const gc_scoring_t *sc, const gc_params_t *p,
                   const uint32_t *packed_q, const uint32_t *packed_t,
                   const uint32_t *q_lens, const uint32_t *t_lens,
                   const uint32_t *q_offs, const uint32_t *t_offs,
                   const uint32_t *seed_scores, uint32_t n, gc_results_t *res, int n_threads)
{
    uint32_t max_q8 = 0, max_t8 = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t a = n_words(q_lens[i]) * 8, b = n_words(t_lens[i]) * 8;
        if (a > max_q8) max_q8 = a;
        if (b > max_t8) max_t8 = b;
    }
    uint32_t mql = p->max_query_len;
    if (mql == 0) { mql = max_q8 > max_t8 ? max_q8 : max_t8; mql += 16; }
    gc_uint4 *tb = NULL;
    if (p->start_mode == GC_WITH_TB && (p->algo == GC_LOCAL || p->algo == GC_GLOBAL || p->algo == GC_BANDED)) {
        uint64_t tiles = ((uint64_t)max_q8 * (max_t8 + 8) + 31) / 32 + 2;                 /* ctors.cpp:112-115 (+slack for Q5 reads) */
        tb = (gc_uint4 *)calloc(tiles * n, sizeof(gc_uint4));
        if (!tb) return -1;
    }
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
    int err = 0;
#pragma omp parallel
    {
        colstate_t *cs = (colstate_t *)malloc(sizeof(colstate_t) * (mql + 16));
        uint32_t *rq = (uint32_t *)malloc(sizeof(uint32_t) * ((mql >> 3) + 2));
        uint32_t *rt = (uint32_t *)malloc(sizeof(uint32_t) * ((mql >> 3) + 2));
        int32_t *eh_h = (int32_t *)malloc(sizeof(int32_t) * (mql + 16));
        int32_t *eh_e = (int32_t *)malloc(sizeof(int32_t) * (mql + 16));
        if (!cs || !rq || !rt || !eh_h || !eh_e) err = -1;
        else {
#pragma omp for schedule(dynamic, 16)
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t *qw = packed_q + (q_offs[i] >> 3), *tw = packed_t + (t_offs[i] >> 3);
                gc_result_t r; memset(&r, 0, sizeof(r));
                switch (p->algo) {
                case GC_LOCAL:
                    local_one(sc, qw, tw, q_lens[i], t_lens[i], p->start_mode, p->second_best, cs, tb, n, i, &r);
                    break;
                case GC_SEMI_GLOBAL:
                    semiglobal_one(sc, qw, tw, q_lens[i], t_lens[i], p->start_mode, p->second_best, p->head, p->tail, mql, cs, rq, rt, &r);
                    break;
                case GC_GLOBAL:
                    global_one(sc, qw, tw, q_lens[i], t_lens[i], p->start_mode, cs, tb, n, i, &r);
                    break;
                case GC_BANDED:
                    if (p->start_mode == GC_WITH_TB) banded_tb_one(sc, qw, tw, q_lens[i], t_lens[i], p->k_band, cs, tb, n, i, &r);
                    else banded_one(sc, qw, tw, q_lens[i], t_lens[i], p->k_band, cs, &r);
                    break;
                case GC_KSW:
                    ksw_one(sc, qw, tw, q_lens[i], t_lens[i], seed_scores ? seed_scores[i] : 0, eh_h, eh_e, mql + 16, &r);
                    break;
                default: break;
                }
                if (tb && res->cigar && res->n_cigar_ops)
                    res->n_cigar_ops[i] = get_tb_one(sc, p->algo != GC_GLOBAL, tb, n, i, q_lens[i], t_lens[i], &r, res->cigar + q_offs[i]);
                if (res->aln_score) res->aln_score[i] = r.score;
                if (res->query_batch_end) res->query_batch_end[i] = r.q_end;
                if (res->target_batch_end) res->target_batch_end[i] = r.t_end;
                if (res->query_batch_start) res->query_batch_start[i] = r.q_start;
                if (res->target_batch_start) res->target_batch_start[i] = r.t_start;
                if (res->aln_score2) res->aln_score2[i] = r.score2;
                if (res->query_batch_end2) res->query_batch_end2[i] = r.q_end2;
                if (res->target_batch_end2) res->target_batch_end2[i] = r.t_end2;
            }
        }
        free(cs); free(rq); free(rt); free(eh_h); free(eh_e);
    }
    free(tb);
    return err;
}

/* reverse-complement a whole batch (gasal_align.cu:232-246) */
void gc_revcomp_batch(uint32_t *packed_q, uint32_t *packed_t, const uint32_t *q_lens, const uint32_t *t_lens,
                      const uint32_t *q_offs, const uint32_t *t_offs, const uint8_t *q_op, const uint8_t *t_op, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        if (q_op[i] == 0 && t_op[i] == 0) continue;
        gc_revcomp_one(packed_q + (q_offs[i] >> 3), q_lens[i], q_op[i]);
        gc_revcomp_one(packed_t + (t_offs[i] >> 3), t_lens[i], t_op[i]);
    }
}
