"""GPU parity: the CUDA engine (through the C ABI) must be BIT-EXACT against the oracle on every
result field, for every mode of the GASAL2 path (a11-a15, Q1-Q14)."""
import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
from oracle import pyoracle as po

from helpers import K3, K5, K2ND, assert_same, oracle_align, cigar_slices, cigar_rescore

pytestmark = pytest.mark.gpu

SCORINGS = [(1, 4, 6, 1), (2, 3, 5, 2), (5, 4, 10, 1)]


def ragged(seed, n=600, **kw):
    return synth.random_ragged_pairs(n, 9, 220, 9, 330, seed=seed, n_frac=0.01, **kw)


def run(batch, *, policy=0, **kw):
    lib = capi.load()
    lib.gasalx_set_kernel_policy(policy)
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
        res = al.align(batch)
        res["_kernel"] = al.kernel_name(0)
    lib.gasalx_set_kernel_policy(0)
    return res


@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("policy", [0, 1])
def test_local_score_end(oracle, scoring, policy):
    b = ragged(11)
    want = oracle_align(oracle, b, algo=po.LOCAL, scoring=scoring)
    got = run(b, policy=policy, algo=capi.LOCAL, scoring=scoring)
    assert_same(got, want, K3, f"local {scoring} policy={policy} kernel={got['_kernel']}")


@pytest.mark.parametrize("policy", [0, 1])
def test_local_with_start(oracle, policy):
    b = ragged(12)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START)
    got = run(b, policy=policy, algo=capi.LOCAL, start=capi.WITH_START)
    assert_same(got, want, K5, f"local start kernel={got['_kernel']}")


def test_local_second_best(oracle):
    b = ragged(13)
    want = oracle_align(oracle, b, algo=po.LOCAL, second_best=True)
    got = run(b, algo=capi.LOCAL, second_best=True)
    assert_same(got, want, K3 + K2ND, "local second-best")


@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("policy,kernel", [(0, "s16x2"), (1, "int32")])
def test_local_traceback(oracle, scoring, policy, kernel):
    """policy 0: direction matrix from the packed s16x2 lanes (sign-bit gather); policy 1: the general int32 kernel"""
    b = ragged(14, n=600)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_TB, scoring=scoring)
    got = run(b, policy=policy, algo=capi.LOCAL, start=capi.WITH_TB, scoring=scoring)
    assert got["_kernel"] == kernel
    assert_same(got, want, K5 + ["n_cigar_ops"], f"local tb {kernel}")
    gc, wc = cigar_slices(got, b), cigar_slices(want, b)
    assert gc == wc
    if scoring == (1, 4, 6, 1):
        # property: at the default scoring every CIGAR re-scores to aln_score
        for i, c in enumerate(gc):
            if got["aln_score"][i] > 0:
                assert cigar_rescore(c, scoring) == got["aln_score"][i]


@pytest.mark.parametrize("head", [po.NONE, po.QUERY, po.TARGET, po.BOTH])
@pytest.mark.parametrize("tail", [po.NONE, po.QUERY, po.TARGET, po.BOTH])
def test_semiglobal_all_head_tail(oracle, head, tail):
    b = ragged(15 + 4 * head + tail, n=300)
    mql = 1024
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, head=head, tail=tail, start=po.WITH_START, max_query_len=mql)
    got = run(b, algo=capi.SEMI_GLOBAL, head=head, tail=tail, start=capi.WITH_START)
    assert_same(got, want, K5, f"semi-global h{head} t{tail} with start")
    # every variant runs on the s16x2 lanes: a charged left column = per-lane start values of the sweep, TAIL in {QUERY, BOTH} = per-lane
    # bottom-row trackers, TAIL = NONE = constants
    assert got["_kernel"] == "s16x2", got["_kernel"]
    got = run(b, algo=capi.SEMI_GLOBAL, head=head, tail=tail)
    assert_same(got, want, K3, f"semi-global h{head} t{tail}")
    pol1 = run(b, policy=1, algo=capi.SEMI_GLOBAL, head=head, tail=tail, start=capi.WITH_START)
    assert pol1["_kernel"] == "int32"
    assert_same(pol1, want, K5, f"semi-global h{head} t{tail} with start, int32")
    if True:
        # other scorings, reads with N in 20 % of the pairs (N-aware pass), longer targets (the charged left column reaches further down),
        # queries longer than their targets (the bottom-row scan can land on column ref_len: the reference's `maxXY_x != ref_len` test)
        for seed, scoring in ((3, (2, 3, 5, 2)), (4, (5, 4, 10, 1)), (5, (1, 1, 0, 1))):
            b2 = synth.with_n_bases(synth.random_ragged_pairs(400, 9, 330, 9, 700 if seed != 4 else 120, seed=900 + 16 * seed + 4 * head + tail), 0.2, seed=seed)
            want = oracle_align(oracle, b2, algo=po.SEMI_GLOBAL, head=head, tail=tail, start=po.WITH_START, max_query_len=mql, scoring=scoring)
            got = run(b2, algo=capi.SEMI_GLOBAL, head=head, tail=tail, start=capi.WITH_START, scoring=scoring)
            assert got["_kernel"] == "s16x2"
            assert_same(got, want, K5, f"semi-global h{head} t{tail} {scoring} with start, N reads")


@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("policy", [0, 1])
def test_semiglobal_default_with_start(oracle, scoring, policy):
    b = ragged(40)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START, scoring=scoring, max_query_len=1024)
    got = run(b, policy=policy, algo=capi.SEMI_GLOBAL, start=capi.WITH_START, scoring=scoring)
    assert_same(got, want, K5, f"semi-global default {scoring} kernel={got['_kernel']}")


def test_semiglobal_second_best(oracle):
    b = ragged(41, n=300)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, second_best=True, max_query_len=1024)
    got = run(b, algo=capi.SEMI_GLOBAL, second_best=True)
    assert_same(got, want, K3 + K2ND, "semi-global second-best")


@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("policy", [0, 1])
def test_global_score(oracle, scoring, policy):
    b = ragged(42)
    want = oracle_align(oracle, b, algo=po.GLOBAL, scoring=scoring)
    got = run(b, policy=policy, algo=capi.GLOBAL, scoring=scoring)
    assert_same(got, want, ["aln_score"], f"global {scoring}")


def test_global_traceback(oracle):
    b = ragged(43, n=300)
    want = oracle_align(oracle, b, algo=po.GLOBAL, start=po.WITH_TB)
    got = run(b, algo=capi.GLOBAL, start=capi.WITH_TB)
    assert_same(got, want, ["aln_score"], "global tb score")
    # parity domain: neither length a multiple of 8 (otherwise the walker's first tile is never written)
    ok = (b.q_lens % 8 != 0) & (b.t_lens % 8 != 0)
    assert ok.sum() > 100
    assert np.array_equal(got["n_cigar_ops"][ok], want["n_cigar_ops"][ok])
    gc, wc = cigar_slices(got, b), cigar_slices(want, b)
    for i in np.nonzero(ok)[0]:
        assert gc[i] == wc[i], f"pair {i}"


@pytest.mark.parametrize("k_band", [8, 24, 64, 4096])
def test_banded(oracle, k_band):
    b = ragged(44 + k_band, n=400)
    want = oracle_align(oracle, b, algo=po.BANDED, k_band=k_band)
    got = run(b, algo=capi.BANDED, k_band=k_band)
    assert_same(got, want, K3, f"banded k={k_band}")


def test_reverse_complement(oracle):
    b = synth.random_ragged_pairs(400, 9, 200, 9, 300, seed=77)
    rng = np.random.default_rng(5)
    q_op = rng.integers(0, 4, size=b.n).astype(np.uint8); t_op = rng.integers(0, 4, size=b.n).astype(np.uint8)
    pq = oracle.pack(b.q_bytes); pt = oracle.pack(b.t_bytes)
    rq, rt = oracle.revcomp(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, q_op, t_op)
    want = oracle.align(rq, rt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.LOCAL, start=po.WITH_START)
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    with Aligner(algo=capi.LOCAL, start=capi.WITH_START, reverse_complement=True, max_query_len=mq, max_target_len=mt,
                 max_n_alns=b.n, n_streams=1) as al:
        got = al.align(b, q_op=q_op, t_op=t_op)
    assert_same(got, want, K5, "reverse-complement + local")


def test_prepacked_input(oracle):
    b = ragged(90, n=300)
    pq = oracle.pack(b.q_bytes); pt = oracle.pack(b.t_bytes)
    want = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.LOCAL)
    # isPacked: the caller hands over packed words in the "unpacked" pages (src/ctors.cpp:64-67); byte counts are 1/2 per base,
    # offsets still index bases (kernels use offset >> 3)
    packed = synth.PairBatch(pq.view(np.uint8), pt.view(np.uint8), b.q_offs, b.t_offs, b.q_lens, b.t_lens)
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    with Aligner(algo=capi.LOCAL, is_packed=True, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
        al.fill(0, packed)
        al.launch(0, ((packed.q_bytes.size + 7) // 8) * 8, ((packed.t_bytes.size + 7) // 8) * 8, b.n)
        al.wait(0)
        got = al.results(0, b.n)
    assert_same(got, want, K3, "pre-packed input")


def test_edge_shapes(oracle):
    """tiny, one-word, exact-multiple-of-8 and very unbalanced pairs; identical sequences; no similarity"""
    qs = [b"A", b"ACGTACGT", b"ACGTACGTA", b"ACGT" * 40, b"T" * 33, b"ACGTTGCA" * 3, b"G" * 8, b"ACGTN" * 10, b"NNNNNNNN"]
    ts = [b"A", b"ACGTACGT", b"TTACGTACGTATT", b"ACGT" * 40, b"A" * 300, b"ACGTTGCA" * 30, b"C" * 16, b"ACGTN" * 20, b"NNNNNNNNNNNN"]
    b = synth.batch_from_sequences(qs, ts)
    for algo, keys, extra in [(po.LOCAL, K5, dict(start=po.WITH_START)), (po.SEMI_GLOBAL, K5, dict(start=po.WITH_START, max_query_len=1024)),
                              (po.GLOBAL, ["aln_score"], {}), (po.BANDED, K3, dict(k_band=16))]:
        want = oracle_align(oracle, b, algo=algo, **extra)
        extra.pop("max_query_len", None)
        got = run(b, algo=algo, **extra)
        assert_same(got, want, keys, f"edge shapes algo={algo}")


def test_multi_storage_round_robin(oracle):
    """two storages in flight, as test_prog.cpp does; also exercises regrow of device buffers"""
    b = ragged(91, n=900)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START)
    chunks = [(0, 300), (300, 600), (600, 900)]
    with Aligner(algo=capi.LOCAL, start=capi.WITH_START, max_query_len=64, max_target_len=64, max_n_alns=100, n_streams=2) as al:
        outs = {}
        pending = {}
        for ci, (lo, hi) in enumerate(chunks):
            si = ci % 2
            if si in pending:
                plo, phi, pb = pending.pop(si); al.wait(si); outs[plo] = al.results(si, phi - plo)
            sub = b.slice(lo, hi)
            # host pages were sized for 100 x 64 bases: grow them the way a caller would (fill one by one)
            al.fill_one_by_one(si, [sub.query(i) for i in range(sub.n)], [sub.target(i) for i in range(sub.n)])
            qb, tb = al._last_bytes
            al.launch(si, qb, tb, sub.n)
            pending[si] = (lo, hi, sub)
        for si, (plo, phi, pb) in pending.items():
            al.wait(si); outs[plo] = al.results(si, phi - plo)
    for lo, hi in chunks:
        sub_want = {k: want[k][lo:hi] for k in K5}
        assert_same(outs[lo], sub_want, K5, f"chunk {lo}:{hi}")


def run_ksw(batch, seeds, policy=0, **kw):
    lib = capi.load()
    lib.gasalx_set_kernel_policy(policy)
    try:
        mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
        with Aligner(algo=capi.KSW, max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
            res = al.align(batch, seeds=seeds)
            res["_kernel"] = al.kernel_name(0)
    finally:
        lib.gasalx_set_kernel_policy(0)
    return res


@pytest.mark.parametrize("scoring", SCORINGS)
def test_ksw_seed_extension(oracle, scoring):
    """algo=KSW (BWA ksw_extend, ksw_kernel_template.h:46-199): h0 from host_seed_scores, adaptive beg/end window, clip rule.
    Lengths include multiples of 8 (the reference's (qlen >> 3) + 1 word count makes the to-end score unconditional there)."""
    b = synth.random_ragged_pairs(600, 8, 200, 8, 260, seed=95, n_frac=0.01)
    rng = np.random.default_rng(7)
    seeds = rng.integers(0, 60, size=b.n).astype(np.uint32)
    seeds[::7] = 0
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    want = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.KSW, scoring=scoring, seed_scores=seeds, max_query_len=1024)
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    with Aligner(algo=capi.KSW, scoring=scoring, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
        got = al.align(b, seeds=seeds)
        assert al.kernel_name(0) == "int32-smem"           # the (H, E) row as u16 pairs in shared memory (k_ksw_smem)
    assert_same(got, want, K3, f"ksw {scoring}")
    pol1 = run_ksw(b, seeds, policy=1, scoring=scoring)
    assert pol1["_kernel"] == "int32"                      # the row in the global scratch (k_ksw_generic)
    assert_same(pol1, want, K3, f"ksw {scoring}, global row")
    # a seed score that does not fit the u16 row sends the batch to the global-row kernel by itself
    big = seeds.copy(); big[5] = 70000
    want = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.KSW, scoring=scoring, seed_scores=big, max_query_len=1024)
    got = run_ksw(b, big, scoring=scoring)
    assert got["_kernel"] == "int32"
    assert_same(got, want, K3, f"ksw {scoring}, large seed score")


def test_ksw_extension_from_matching_prefix(oracle):
    """extensions that run to the end of the query (gscore branch) and ones that are clipped"""
    rng = np.random.default_rng(3)
    qs, ts, seeds = [], [], []
    for k in range(300):
        L = int(rng.integers(16, 150))
        core = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
        t = np.concatenate([core, rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(0, 40)))])
        q = core.copy()
        if k % 3 == 0:
            q[L // 2:] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L - L // 2)     # second half unrelated: clipped
        qs.append(q.tobytes()); ts.append(t.tobytes()); seeds.append(int(rng.integers(10, 40)))
    b = synth.batch_from_sequences(qs, ts)
    seeds = np.asarray(seeds, dtype=np.uint32)
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    want = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.KSW, seed_scores=seeds, max_query_len=1024)
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    with Aligner(algo=capi.KSW, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
        got = al.align(b, seeds=seeds)
        assert al.kernel_name(0) == "int32-smem"
    assert_same(got, want, K3, "ksw prefix extension")
    assert_same(run_ksw(b, seeds, policy=1), want, K3, "ksw prefix extension, global row")
    # the bench shape: 20,000 extensions of 150 bp reads along 200 bp windows, both kernels
    bk = synth.make_extension_pairs(20_000, 150, 200)
    sk = np.full(bk.n, 20, dtype=np.uint32)
    wk = oracle.align(oracle.pack(bk.q_bytes), oracle.pack(bk.t_bytes), bk.q_lens, bk.t_lens, bk.q_offs, bk.t_offs, algo=po.KSW, seed_scores=sk, max_query_len=1024)
    assert_same(run_ksw(bk, sk), wk, K3, "ksw 150 x 200 extensions")
    assert_same(run_ksw(bk, sk, policy=1), wk, K3, "ksw 150 x 200 extensions, global row")
    assert (got["query_batch_end"] == b.q_lens.astype(np.int32)).sum() > 100      # the to-end branch was exercised


# gap costs that stress the biased frame of the semi-global / global kernels ((row + column) * ext folded into every value):
# no extension cost at all, extension as expensive as opening, free opening, large extension, and a scoring whose
# substitution bytes (+ 2 ext) no longer fit the packed tables so that the engine must fall back by itself
ODD_SCORINGS = [(1, 4, 6, 0), (2, 3, 0, 2), (1, 1, 0, 1), (3, 5, 9, 7), (10, 9, 20, 5), (100, 60, 7, 20)]


@pytest.mark.parametrize("scoring", ODD_SCORINGS)
@pytest.mark.parametrize("policy", [0, 2])
def test_biased_frame_scorings(oracle, scoring, policy):
    """policy 0: s16x2 fast path where eligible; policy 2: the wavefront kernel wherever it is eligible"""
    b = ragged(70 + scoring[3], n=400)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START, scoring=scoring, max_query_len=1024)
    got = run(b, policy=policy, algo=capi.SEMI_GLOBAL, start=capi.WITH_START, scoring=scoring)
    assert_same(got, want, K5, f"semi-global {scoring} policy={policy} kernel={got['_kernel']}")
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, head=po.BOTH, tail=po.TARGET, scoring=scoring, max_query_len=1024)
    got = run(b, policy=policy, algo=capi.SEMI_GLOBAL, head=capi.BOTH, tail=capi.TARGET, scoring=scoring)
    assert_same(got, want, K3, f"semi-global BOTH/TARGET {scoring} policy={policy} kernel={got['_kernel']}")
    want = oracle_align(oracle, b, algo=po.GLOBAL, scoring=scoring)
    got = run(b, policy=policy, algo=capi.GLOBAL, scoring=scoring)
    assert_same(got, want, ["aln_score"], f"global {scoring} policy={policy} kernel={got['_kernel']}")


def test_non_fatal_error_mode(oracle):
    """gasalx_try_aln_async returns the reference's fatal argument errors as a status; the storage stays usable"""
    import ctypes as C
    b = ragged(95, n=200)
    lib = capi.load()
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    with Aligner(algo=capi.LOCAL, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
        al.fill(0, b)
        s = al.vec.a[0]
        rc = lib.gasalx_try_aln_async(C.byref(s), b.q_bytes.size, b.t_bytes.size, s.host_max_n_alns + 1, C.byref(al.params))
        assert rc == 1 and b"actual_n_alns" in lib.gasalx_last_error()
        rc = lib.gasalx_try_aln_async(C.byref(s), b.q_bytes.size + 4, b.t_bytes.size, b.n, C.byref(al.params))
        assert rc == 1 and b"multiple of 8" in lib.gasalx_last_error()
        assert al.done(0) == -2                                      # nothing was launched
        rc = lib.gasalx_try_aln_async(C.byref(s), b.q_bytes.size, b.t_bytes.size, b.n, C.byref(al.params))
        assert rc == 0
        al.wait(0)
        got = al.results(0, b.n)
    assert_same(got, oracle_align(oracle, b, algo=po.LOCAL), K3, "after soft errors")


def test_opt_in_coordinate_fix(oracle):
    """gasalx_set_coordinate_fix(1): un-mirrors LOCAL WITH_START's target_batch_start (reported = 16*floor(ts/8) + 14 - ts, so the
    repaired value must equal the TRUE start, which LOCAL WITH_TB reports) and gives the default semi-global mode meaningful query
    coordinates (Q8); everything else stays bit-exact with the reference, and the default mode (0) is untouched."""
    b = ragged(31)
    lib = capi.load()
    ref_start = run(b, algo=capi.LOCAL, start=capi.WITH_START)
    tb = run(b, algo=capi.LOCAL, start=capi.WITH_TB)                       # true starts from the traceback walker
    lib.gasalx_set_coordinate_fix(1)
    try:
        fixed = run(b, algo=capi.LOCAL, start=capi.WITH_START)
        sg = run(b, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
    finally:
        lib.gasalx_set_coordinate_fix(0)
    hit = ref_start["aln_score"] > 0
    ts = fixed["target_batch_start"]
    assert np.array_equal(16 * (ts // 8) + 14 - ts, ref_start["target_batch_start"])                       # the documented mirror formula
    # = a true start: on the same diagonal as the traceback's start for nearly every pair (co-optimal alignments may begin a few
    # zero-sum bases earlier or later along that diagonal); the mirrored value is on that diagonal for ~1 pair in 8 only
    same_diag = lambda t_start: np.mean((t_start - fixed["query_batch_start"])[hit] == (tb["target_batch_start"] - tb["query_batch_start"])[hit])
    assert same_diag(ts) > 0.9 and same_diag(ref_start["target_batch_start"]) < 0.3
    for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start"):
        assert np.array_equal(fixed[k], ref_start[k]), k
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START)
    for k in ("aln_score", "target_batch_end", "target_batch_start"):
        assert np.array_equal(sg[k], want[k]), k
    assert np.array_equal(sg["query_batch_end"], b.q_lens.astype(np.int32) - 1) and not sg["query_batch_start"].any()
    again = run(b, algo=capi.LOCAL, start=capi.WITH_START)
    assert np.array_equal(again["target_batch_start"], ref_start["target_batch_start"])


def test_banded_score_runs_on_the_packed_kernel(oracle):
    """a14 on the s16x2 lanes (k_banded_tb_s16<NMASK, TB = false>, H-based recurrence of banded.h:99-114 over the tile band): C4's
    shape with N in 10 % of the pairs (N-aware pass), odd band geometries (band wider than the matrix, query longer than the
    target, band of one tile), other scorings, and start_pos = WITH_START (the banded kernel ignores it: gasal_align.h:78-81)"""
    b = synth.with_n_bases(synth.config4(8192, seed=46), 0.1, seed=4)
    want = oracle_align(oracle, b, algo=po.BANDED, k_band=64)
    got = run(b, algo=capi.BANDED, k_band=64)
    assert got["_kernel"] == "s16x2"
    assert_same(got, want, K3, "banded s16x2, C4 shape with N")
    pol1 = run(b, policy=1, algo=capi.BANDED, k_band=64)
    assert pol1["_kernel"] == "int32"
    assert_same(pol1, want, K3, "banded int32, C4 shape with N")
    for k_band, qmax, tmax, scoring in ((8, 40, 330, (1, 4, 6, 1)), (24, 220, 60, (2, 3, 5, 2)), (40, 330, 330, (5, 4, 10, 1)),
                                        (4096, 100, 200, (1, 4, 6, 1)), (16, 300, 300, (1, 1, 0, 1)), (72, 250, 250, (3, 5, 0, 0))):
        b = synth.random_ragged_pairs(700, 9, qmax, 9, tmax, seed=700 + k_band, n_frac=0.02)
        want = oracle_align(oracle, b, algo=po.BANDED, k_band=k_band, scoring=scoring)
        got = run(b, algo=capi.BANDED, k_band=k_band, scoring=scoring)
        assert got["_kernel"] == "s16x2", (k_band, got["_kernel"])
        assert_same(got, want, K3, f"banded s16x2 k={k_band} q<={qmax} t<={tmax} {scoring}")
    b = synth.random_ragged_pairs(300, 9, 200, 9, 300, seed=733)
    want = oracle_align(oracle, b, algo=po.BANDED, k_band=32)
    got = run(b, algo=capi.BANDED, k_band=32, start=capi.WITH_START)
    assert_same(got, want, K3, "banded s16x2 with start_pos = WITH_START")
