"""Parity at BASELINE.json's full sizes: direct oracle comparison where the CPU oracle still finishes in seconds, and
size-independent properties (batch-composition independence, kernel-family agreement, pass consistency) on the big batches."""
import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
from oracle import pyoracle as po

from helpers import K3, K5, assert_same, oracle_align

pytestmark = pytest.mark.gpu


def run(batch, *, policy=0, **kw):
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


def test_config3_full_length_global_vs_oracle(oracle):
    """C3: Needleman-Wunsch 16,384 x 16,384 (NW benchmark input model), full length, 6 pairs (1.6e9 cells on the CPU oracle)"""
    b = synth.config3(6, 16384, seed=3)
    want = oracle_align(oracle, b, algo=po.GLOBAL, max_query_len=16400)
    got = run(b, algo=capi.GLOBAL)
    assert got["_kernel"] == "wave"
    assert_same(got, want, ["aln_score"], "C3 full length")
    assert (np.abs(want["aln_score"]) < 32767).all()          # inside the parity domain


def test_config5_longest_pairs_local_with_start_vs_oracle(oracle):
    """C5: the long end of the 100 bp - 10 kb mix (8 % subs + 4 % indels), local WITH_START, both wavefront lane widths"""
    b = synth.config5(10, seed=55, lo=7000, hi=10000)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START, max_query_len=16400)
    for policy, kernel in ((4, "wave16"), (3, "wave")):
        got = run(b, policy=policy, algo=capi.LOCAL, start=capi.WITH_START)
        assert got["_kernel"] == kernel
        assert_same(got, want, K5, f"C5 long pairs {kernel}")
    assert want["aln_score"].min() > 1000


def test_config2_full_batch_properties():
    """C2 at a full GPU-filling batch (262,144 pairs of 150 x 600): results do not depend on how pairs are batched, the fast path
    agrees with the general kernel on a sample, and the WITH_START launch reproduces the score/end of the plain launch."""
    n = 262_144
    b = synth.config2(n, seed=2)
    for algo in (capi.LOCAL, capi.SEMI_GLOBAL):
        full = run(b, algo=algo, start=capi.WITH_START)
        assert full["_kernel"] == "s16x2"
        plain = run(b, algo=algo)
        assert_same(plain, full, K3, f"algo {algo}: score/end with and without the start pass")
        lo, hi = 100_003, 100_003 + 20_011                       # an unaligned slice: different CTA / warp neighbours
        sub = b.slice(lo, hi)
        part = run(sub, algo=algo, start=capi.WITH_START)
        assert_same(part, {k: full[k][lo:hi] for k in K5}, K5, f"algo {algo}: batch-composition independence")
        gen = run(sub, policy=1, algo=algo, start=capi.WITH_START)
        assert gen["_kernel"] == "int32"
        assert_same(part, gen, K5, f"algo {algo}: fast path vs general kernel")
        # domain sanity: every read was cut from its window with 2 % substitutions, so it aligns almost end to end
        assert np.median(full["aln_score"]) > 100
        if algo == capi.LOCAL:
            assert (full["query_batch_end"] < 152).all() and (full["target_batch_end"] < 600).all()


def test_config4_full_shape_banded_cigar_properties(oracle):
    """C4 shape at 65,536 pairs: banded + CIGAR equals unbanded local + CIGAR from the engine itself (the oracle-pinned equality on a
    3,000-pair sample is in test_banded_tb_gpu.py), and a sample re-scores"""
    from helpers import cigar_slices, cigar_rescore
    b = synth.config4(65_536, seed=44)
    ban = run(b, algo=capi.BANDED, start=capi.WITH_TB, k_band=64)
    loc = run(b, algo=capi.LOCAL, start=capi.WITH_TB)
    assert_same(ban, loc, K5 + ["n_cigar_ops"], "banded k=64 + CIGAR vs unbanded")
    sub = b.slice(0, 2000)
    gc = cigar_slices({k: ban[k][:2000] if k != "cigar" else ban[k] for k in ("n_cigar_ops", "cigar")}, sub)
    lc = cigar_slices({k: loc[k][:2000] if k != "cigar" else loc[k] for k in ("n_cigar_ops", "cigar")}, sub)
    assert gc == lc
    for i, c in enumerate(gc):
        assert cigar_rescore(c, (1, 4, 6, 1)) == ban["aln_score"][i]


def test_config4_full_batch_banded_score(oracle):
    """C4 at a full 262,144-pair batch, score/end only: the s16x2 banded sweep (banded.h's H-based recurrence over the tile band) equals
    the int32 banded kernel on every pair, the oracle on 20,000 of them, and results do not depend on how the batch was cut"""
    b = synth.config4(262_144, seed=45)
    full = run(b, algo=capi.BANDED, k_band=64)
    assert full["_kernel"] == "s16x2"
    gen = run(b, policy=1, algo=capi.BANDED, k_band=64)
    assert gen["_kernel"] == "int32"
    assert_same(full, gen, K3, "banded score/end: s16x2 vs int32, 262144 pairs")
    sub = b.slice(100_000, 120_000)
    want = oracle_align(oracle, sub, algo=po.BANDED, k_band=64)
    assert_same({k: full[k][100_000:120_000] for k in K3}, want, K3, "banded score/end vs the oracle, 20000 pairs")
    part = run(sub, algo=capi.BANDED, k_band=64)
    assert_same(part, want, K3, "banded score/end: batch-composition independence")
    assert np.median(full["aln_score"]) > 150


def test_sw_benchmark_full_batch_two_packed_kernels():
    """the nvbio adapter at sw-benchmark's batch size (262,144 reads of 150 bp against 600 bp windows): the quad-strip kernel and the
    band-blocked kernel -- two independent implementations (K2's blocking / the reference's blocking) -- agree on every score of all
    three alignment types and on the GLOBAL / SEMI_GLOBAL sinks, and with the int32 kernel on a 30,000-job slice"""
    from genomics_gpu_b200 import nvb
    pats, poffs, texts, toffs = synth.nvb_jobs(262_144, 150, 150, 600, 600, seed=71, sub=0.03, indel=0.005)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    k = 30_000
    small = nvb.Batch(pats[:poffs[k]], poffs[:k + 1], texts[:toffs[k]], toffs[:k + 1])
    try:
        with nvb.Stream(batch) as st, nvb.Stream(small) as st_small:
            for type_ in (nvb.GLOBAL, nvb.LOCAL, nvb.SEMI_GLOBAL):
                al = nvb.aligner(nvb.GOTOH, type_, 2, -1, -2, -1)
                sinks = type_ != nvb.LOCAL
                nvb.set_policy(0); st.enact(al, sinks); a = st.results(sinks); ka = st.kernel()
                nvb.set_policy(2); st.enact(al, sinks); b2 = st.results(sinks); kb = st.kernel()
                assert ka == "s16x2-strip" and kb == "s16x2", (ka, kb)
                assert np.array_equal(a[0], b2[0]), type_
                if sinks:
                    assert np.array_equal(a[1], b2[1]) and np.array_equal(a[2], b2[2]), type_
                nvb.set_policy(1); st_small.enact(al, sinks); c = st_small.results(sinks)
                assert st_small.kernel() == "int32" and np.array_equal(c[0], a[0][:k]), type_
    finally:
        nvb.set_policy(0)
