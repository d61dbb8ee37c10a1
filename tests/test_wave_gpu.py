"""GPU parity of the intra-task wavefront kernel (K3, dp_wave.cuh) against the oracle, through the C ABI.

policy 2 forces the wavefront kernel wherever it is eligible, so the ragged short batches of the general parity suite
exercise every edge (1-strip targets, partial row-blocks, pads, N bases, exotic IUPAC codes); the long batches exercise the
multi-block pipeline (boundary hand-over through global memory, tickets, per-pair reduction)."""
import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
from oracle import pyoracle as po

from helpers import K3, K5, assert_same, oracle_align

pytestmark = pytest.mark.gpu

SCORINGS = [(1, 4, 6, 1), (2, 3, 5, 2), (5, 4, 10, 1)]


def run(batch, *, policy=2, **kw):
    lib = capi.load()
    lib.gasalx_set_kernel_policy(policy)
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    try:
        with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
            res = al.align(batch)
            res["_kernel"] = al.kernel_name(0)
    finally:
        lib.gasalx_set_kernel_policy(0)
    return res


def ragged(seed, n=500, **kw):
    return synth.random_ragged_pairs(n, 9, 220, 9, 700, seed=seed, n_frac=0.01, **kw)


def long_pairs(seed, n=24, lo=1500, hi=5000):
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for _ in range(n):
        L = int(rng.integers(lo, hi))
        t, q = synth.nw_style_pair(rng, L)
        qs.append(q); ts.append(t)
    return synth.batch_from_sequences(qs, ts)


@pytest.mark.parametrize("scoring", SCORINGS)
@pytest.mark.parametrize("policy,kernel", [(4, "wave16"), (3, "wave")])
def test_wave_local_ragged(oracle, scoring, policy, kernel):
    """policy 4: s16x2 wavefront kernel (k_wave16_local); policy 3: the int32 wavefront kernel"""
    b = ragged(301)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START, scoring=scoring)
    got = run(b, policy=policy, algo=capi.LOCAL, start=capi.WITH_START, scoring=scoring)
    assert got["_kernel"] == kernel
    assert_same(got, want, K5, f"{kernel} local {scoring}")
    got = run(b, policy=policy, algo=capi.LOCAL, scoring=scoring)
    assert_same(got, want, K3, f"{kernel} local score/end {scoring}")


@pytest.mark.parametrize("scoring", SCORINGS)
def test_wave_global_ragged(oracle, scoring):
    b = ragged(302)
    want = oracle_align(oracle, b, algo=po.GLOBAL, scoring=scoring)
    got = run(b, policy=3, algo=capi.GLOBAL, scoring=scoring)            # policy 3: int32 lanes
    assert got["_kernel"] == "wave"
    assert_same(got, want, ["aln_score"], f"wave global {scoring}")


@pytest.mark.parametrize("head", [po.TARGET, po.BOTH])
@pytest.mark.parametrize("scoring", SCORINGS[:2])
def test_wave_semiglobal_ragged(oracle, scoring, head):
    b = ragged(303 + head)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, head=head, tail=po.TARGET, start=po.WITH_START, scoring=scoring, max_query_len=1024)
    got = run(b, policy=3, algo=capi.SEMI_GLOBAL, head=head, tail=capi.TARGET, start=capi.WITH_START, scoring=scoring)   # policy 3: int32 lanes
    assert got["_kernel"] == "wave"
    assert_same(got, want, K5, f"wave semi-global head={head} {scoring}")


def test_wave_exotic_codes(oracle):
    """IUPAC letters other than ACGTN are ordinary symbols (equal -> match); they take the compare path"""
    rng = np.random.default_rng(9)
    alpha = np.frombuffer(b"ACGTACGTACGTNRYK", dtype=np.uint8)
    qs = [bytes(alpha[rng.integers(0, alpha.size, size=int(rng.integers(20, 400)))]) for _ in range(200)]
    ts = [bytes(alpha[rng.integers(0, alpha.size, size=int(rng.integers(20, 600)))]) for _ in range(200)]
    b = synth.batch_from_sequences(qs, ts)
    for algo, keys, extra in [(po.LOCAL, K5, dict(start=po.WITH_START)), (po.GLOBAL, ["aln_score"], {}),
                              (po.SEMI_GLOBAL, K5, dict(start=po.WITH_START, max_query_len=1024))]:
        want = oracle_align(oracle, b, algo=algo, **extra)
        extra.pop("max_query_len", None)
        for policy in ((4, 3) if algo in (po.LOCAL, po.GLOBAL) else (2,)):
            got = run(b, policy=policy, algo=algo, **extra)
            assert got["_kernel"] in ("wave", "wave16")
            assert_same(got, want, keys, f"wave exotic codes algo={algo} policy={policy}")


def test_wave_long_local_with_start(oracle):
    b = long_pairs(41)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START, max_query_len=8192)
    got = run(b, policy=0, algo=capi.LOCAL, start=capi.WITH_START)          # policy 0: long batches pick the wavefront kernel themselves
    assert got["_kernel"] in ("wave", "wave16")
    assert_same(got, want, K5, "wave long local")
    for policy, kernel in ((4, "wave16"), (3, "wave")):
        got = run(b, policy=policy, algo=capi.LOCAL, start=capi.WITH_START)
        assert got["_kernel"] == kernel
        assert_same(got, want, K5, f"{kernel} long local")


def test_wave_long_global(oracle):
    b = long_pairs(42)
    want = oracle_align(oracle, b, algo=po.GLOBAL, max_query_len=8192)
    for policy, kernel in ((3, "wave"), (4, "wave16")):
        got = run(b, policy=policy, algo=capi.GLOBAL)
        assert got["_kernel"] == kernel
        assert_same(got, want, ["aln_score"], f"{kernel} long global")


@pytest.mark.parametrize("scoring", SCORINGS + [(1, 4, 6, 0), (2, 3, 0, 2), (3, 5, 9, 7)])
def test_wave16_global_ragged(oracle, scoring):
    """policy 4: k_wave16_global (s16x2 lanes, biased + shifted frame) on ragged short pairs: every boundary case of the
    half-entry logic (pairs shorter than one strip, one lane, one row-block), N codes (ordinary symbols under the global rule)"""
    b = ragged(60 + scoring[0])
    want = oracle_align(oracle, b, algo=po.GLOBAL, scoring=scoring)
    got = run(b, policy=4, algo=capi.GLOBAL, scoring=scoring)
    assert got["_kernel"] == "wave16"
    assert_same(got, want, ["aln_score"], f"wave16 global {scoring}")


def test_wave_long_semiglobal_with_start(oracle):
    rng = np.random.default_rng(43)
    qs, ts = [], []
    for _ in range(24):
        L = int(rng.integers(800, 3000))
        t, q = synth.nw_style_pair(rng, L)
        flank_l = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(0, 900))))
        flank_r = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(0, 900))))
        qs.append(q); ts.append(flank_l + t + flank_r)
    b = synth.batch_from_sequences(qs, ts)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=8192)
    got = run(b, policy=0, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
    assert got["_kernel"] == "wave"
    assert_same(got, want, K5, "wave long semi-global")


@pytest.mark.parametrize("head", [po.TARGET, po.BOTH])
@pytest.mark.parametrize("scoring", SCORINGS + [(1, 4, 6, 0), (2, 3, 0, 2), (3, 5, 9, 7)])
def test_wave16_semiglobal_ragged(oracle, scoring, head):
    """policy 4: k_wave16_semi (s16x2 lanes, biased + shifted frame, free left column) on ragged short pairs: pairs shorter than one
    strip / one lane / one row-block, N codes, zero and large extension costs, both HEAD variants, with and without the start pass"""
    b = ragged(70 + scoring[0] + 7 * head)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, head=head, tail=po.TARGET, start=po.WITH_START, scoring=scoring, max_query_len=1024)
    got = run(b, policy=4, algo=capi.SEMI_GLOBAL, head=head, tail=capi.TARGET, start=capi.WITH_START, scoring=scoring)
    assert got["_kernel"] == "wave16"
    assert_same(got, want, K5, f"wave16 semi-global head={head} {scoring}")
    got = run(b, policy=4, algo=capi.SEMI_GLOBAL, head=head, tail=capi.TARGET, scoring=scoring)
    assert got["_kernel"] == "wave16"
    assert_same(got, want, K3, f"wave16 semi-global (no start) head={head} {scoring}")


def test_wave16_semiglobal_long_and_automatic(oracle):
    """long reads in flanked windows on the s16x2 lanes (forced), then enough of them that the launcher picks those lanes itself;
    IUPAC codes take the compare path"""
    rng = np.random.default_rng(44)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def make(n, lo, hi, flank):
        qs, ts = [], []
        for _ in range(n):
            t, q = synth.nw_style_pair(rng, int(rng.integers(lo, hi)))
            qs.append(q); ts.append(bytes(rng.choice(acgt, size=int(rng.integers(0, flank)))) + t + bytes(rng.choice(acgt, size=int(rng.integers(0, flank)))))
        return synth.batch_from_sequences(qs, ts)

    b = make(24, 800, 3000, 900)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=8192)
    for policy, kernel in ((4, "wave16"), (3, "wave")):
        got = run(b, policy=policy, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
        assert got["_kernel"] == kernel
        assert_same(got, want, K5, f"{kernel} long semi-global")
    b = make(700, 700, 1500, 500)                                          # >= 444 row-blocks of 512 rows: automatic choice
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=8192)
    got = run(b, policy=0, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
    assert got["_kernel"] == "wave16"
    assert_same(got, want, K5, "wave16 semi-global, launcher's own choice")
    alpha = np.frombuffer(b"ACGTACGTACGTNRYK", dtype=np.uint8)
    qs = [bytes(alpha[rng.integers(0, alpha.size, size=int(rng.integers(20, 400)))]) for _ in range(200)]
    ts = [bytes(alpha[rng.integers(0, alpha.size, size=int(rng.integers(20, 600)))]) for _ in range(200)]
    b = synth.batch_from_sequences(qs, ts)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=1024)
    got = run(b, policy=4, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
    assert got["_kernel"] == "wave16"
    assert_same(got, want, K5, "wave16 semi-global, IUPAC codes")


def test_wave_matches_general_kernel_on_mixed_lengths(oracle):
    """size-independent property: the wavefront kernel and the one-thread-per-pair kernel agree on a mixed 100 bp - 6 kb batch"""
    b = synth.config5(96, seed=7, lo=100, hi=6000)
    a = run(b, policy=1, algo=capi.LOCAL, start=capi.WITH_START)
    w = run(b, policy=0, algo=capi.LOCAL, start=capi.WITH_START)
    # policy 0 regroups the batch by length class: short queries on the s16x2 fast path, long ones on the wavefront kernel
    assert a["_kernel"] == "int32" and (w["_kernel"] in ("wave", "wave16") or w["_kernel"].startswith("mixed(s16x2+wave"))
    w2 = run(b, policy=2, algo=capi.LOCAL, start=capi.WITH_START)
    assert w2["_kernel"] in ("wave", "wave16")
    assert_same(w2, a, K5, "wavefront kernel alone vs general kernel, mixed lengths")
    assert_same(w, a, K5, "wave vs general kernel, mixed lengths")


@pytest.mark.parametrize("scoring", [(1, 4, 6, 1), (2, 3, 5, 2)])
def test_wave16_local_traceback_long_pairs(oracle, scoring):
    """LOCAL WITH_TB on long pairs (missing #3): the s16x2 wavefront kernel emits the direction matrix into per-pair
    regions and k_get_tb walks them -- no fall-back to the one-thread-per-pair int32 kernel.  CIGAR bytes, starts, ends vs the oracle."""
    from helpers import cigar_slices, cigar_rescore
    b = long_pairs(410 + scoring[0], n=20, lo=400, hi=4000)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_TB, scoring=scoring, max_query_len=8192)
    got = run(b, policy=0, algo=capi.LOCAL, start=capi.WITH_TB, scoring=scoring)
    assert got["_kernel"] == "wave16"
    assert_same(got, want, K5 + ["n_cigar_ops"], f"wave16 local WITH_TB {scoring}")
    gc, wc = cigar_slices(got, b), cigar_slices(want, b)
    assert gc == wc
    if scoring == (1, 4, 6, 1):
        for i, c in enumerate(gc):
            assert cigar_rescore(c, scoring) == got["aln_score"][i]


def test_wave16_local_traceback_ragged_short(oracle):
    """the same kernel forced onto ragged short pairs (policy 4): one-strip targets, partial row-blocks, pads, N and IUPAC columns"""
    from helpers import cigar_slices
    b = ragged(415, n=400)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_TB)
    got = run(b, policy=4, algo=capi.LOCAL, start=capi.WITH_TB)
    assert got["_kernel"] == "wave16"
    assert_same(got, want, K5 + ["n_cigar_ops"], "wave16 local WITH_TB ragged")
    assert cigar_slices(got, b) == cigar_slices(want, b)
