"""CPU tests: pin the oracle (oracle/gotoh_compat.c) to the reference.

1. against the committed golden vectors that were produced by the reference's own kernels
   (tests/golden/make_golden.py) -- runs anywhere;
2. live against the reference kernels compiled for the host (oracle/_ref) on fresh seeds -- runs wherever
   that library exists (it is built in the dev container and travels to the GPU box);
3. known-answer probes recorded in SURVEY.md section 8 (pack nibbles, quirks Q3/Q4/Q8).
"""
import os

import numpy as np
import pytest

from genomics_gpu_b200 import synth
from oracle import pyoracle as po

from helpers import K3, K5, K2ND, assert_same

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_synth.npz"))


def gold_batch():
    return synth.PairBatch(GOLD["q_bytes"], GOLD["t_bytes"], GOLD["q_offs"], GOLD["t_offs"], GOLD["q_lens"], GOLD["t_lens"])


def golden_modes():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m.modes()


def test_pack_matches_golden(oracle):
    assert np.array_equal(oracle.pack(GOLD["q_bytes"]), GOLD["packed_q"])
    assert np.array_equal(oracle.pack(GOLD["t_bytes"]), GOLD["packed_t"])


def test_pack_known_answer(oracle):
    # "ACGTNacg" -> 0x1374e137
    assert oracle.pack(np.frombuffer(b"ACGTNacg", dtype=np.uint8))[0] == 0x1374E137


@pytest.mark.parametrize("name,kw,keys", golden_modes(), ids=[m[0] for m in golden_modes()])
def test_oracle_matches_reference_golden(oracle, name, kw, keys):
    b = gold_batch()
    kw = dict(kw)
    if kw["algo"] == po.KSW:
        kw["seed_scores"] = GOLD["seed_scores"]
    got = oracle.align(GOLD["packed_q"], GOLD["packed_t"], b.q_lens, b.t_lens, b.q_offs, b.t_offs, max_query_len=1024, **kw)
    want = {k: GOLD[f"{name}/{k}"] for k in keys}
    if name == "global_tb":
        # parity domain: the walker's first tile is unwritten when a length is a multiple of 8
        ok = (b.q_lens % 8 != 0) & (b.t_lens % 8 != 0)
        assert np.array_equal(got["aln_score"], want["aln_score"])
        assert np.array_equal(got["n_cigar_ops"][ok], want["n_cigar_ops"][ok])
        for i in np.nonzero(ok)[0]:
            o, n = int(b.q_offs[i]), int(want["n_cigar_ops"][i])
            assert np.array_equal(got["cigar"][o:o + n], want["cigar"][o:o + n])
        return
    if "cigar" in keys:
        for i in range(b.n):
            o, n = int(b.q_offs[i]), int(want["n_cigar_ops"][i])
            assert np.array_equal(got["cigar"][o:o + n], want["cigar"][o:o + n]), f"cigar of pair {i}"
        keys = [k for k in keys if k != "cigar"]
    assert_same(got, want, keys, name)


@pytest.mark.skipif(not po.RefHost.available(1024), reason="oracle/_ref host build of the reference not present")
@pytest.mark.parametrize("seed", [101, 202])
def test_oracle_matches_reference_live(oracle, seed):
    ref = po.RefHost(1024)
    b = synth.random_ragged_pairs(300, 9, 260, 9, 400, seed=seed, n_frac=0.02)
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    rq, rt = ref.pack(b.q_bytes, b.t_bytes, b.n)
    assert np.array_equal(pq, rq) and np.array_equal(pt, rt)
    for sc in [(1, 4, 6, 1), (3, 2, 4, 2)]:
        for kw, keys in [(dict(algo=po.LOCAL, start=po.WITH_START), K5), (dict(algo=po.LOCAL, second_best=True), K3 + K2ND),
                         (dict(algo=po.SEMI_GLOBAL, start=po.WITH_START), K5), (dict(algo=po.SEMI_GLOBAL, head=po.QUERY, tail=po.BOTH, start=po.WITH_START), K5),
                         (dict(algo=po.GLOBAL), ["aln_score"]), (dict(algo=po.BANDED, k_band=32), K3)]:
            want = ref.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=sc, **kw)
            got = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=sc, max_query_len=1024, **kw)
            assert_same(got, want, keys, f"{kw} {sc}")


@pytest.mark.skipif(not po.RefHost.available(1024), reason="oracle/_ref host build of the reference not present")
@pytest.mark.parametrize("k_band,qmax,tmax,scoring", [(8, 40, 330, (1, 4, 6, 1)), (24, 220, 60, (2, 3, 5, 2)), (40, 330, 330, (5, 4, 10, 1)),
                                                       (4096, 100, 200, (1, 4, 6, 1)), (16, 300, 300, (1, 1, 0, 1)), (72, 250, 250, (3, 5, 0, 0))])
def test_oracle_matches_reference_banded_shapes(oracle, k_band, qmax, tmax, scoring):
    """gasal_banded_tiled_kernel (banded.h:10-139) on the band geometries and scorings the s16x2 banded sweep is tested with
    (tests/test_parity_gpu.py::test_banded_score_runs_on_the_packed_kernel): band of one tile, band wider than the matrix, query
    longer than the target, free gaps -- the restatement against the reference's own kernel compiled for the host"""
    ref = po.RefHost(1024)
    b = synth.random_ragged_pairs(250, 9, qmax, 9, tmax, seed=700 + k_band, n_frac=0.02)
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    want = ref.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=scoring, algo=po.BANDED, k_band=k_band)
    got = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=scoring, max_query_len=1024, algo=po.BANDED, k_band=k_band)
    assert_same(got, want, K3, f"banded k={k_band} {scoring}")


@pytest.mark.skipif(not po.RefHost.available(1024), reason="oracle/_ref host build of the reference not present")
@pytest.mark.parametrize("head", [po.NONE, po.QUERY, po.TARGET, po.BOTH])
def test_oracle_matches_reference_all_head_tail(oracle, head):
    """all 16 semiglobal_skipping_head / _tail variants (semiglobal_kernel_template.h:39-388) with start positions: the restatement
    against the reference's own kernels compiled for the host, incl. reads with N"""
    ref = po.RefHost(1024)
    b = synth.random_ragged_pairs(160, 9, 200, 9, 330, seed=300 + head, n_frac=0.05)
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    for tail in (po.NONE, po.QUERY, po.TARGET, po.BOTH):
        for sc in [(1, 4, 6, 1), (2, 3, 5, 2)]:
            kw = dict(algo=po.SEMI_GLOBAL, head=head, tail=tail, start=po.WITH_START)
            want = ref.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=sc, **kw)
            got = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=sc, max_query_len=1024, **kw)
            assert_same(got, want, K5, f"head {head} tail {tail} {sc}")


@pytest.mark.skipif(not po.RefHost.available(1024), reason="oracle/_ref host build of the reference not present")
@pytest.mark.parametrize("scoring", [(1, 4, 6, 1), (2, 3, 5, 2), (1, 1, 0, 1), (3, 2, 8, 0)])
def test_oracle_matches_reference_on_overhanging_reads(oracle, scoring):
    """reads that hang off their window / long gaps / zero gap costs: the shapes the GPU start-pass pruning is tested on
    (tests/test_parity_holes_gpu.py) -- here the oracle itself is pinned to the reference's kernels on them"""
    ref = po.RefHost(1024)
    b = synth.overhanging_reads(600, seed=77)
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    for kw in (dict(algo=po.LOCAL, start=po.WITH_START), dict(algo=po.SEMI_GLOBAL, start=po.WITH_START),
               dict(algo=po.SEMI_GLOBAL, head=po.BOTH, start=po.WITH_START)):
        want = ref.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=scoring, **kw)
        got = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, scoring=scoring, max_query_len=1024, **kw)
        assert_same(got, want, K5, f"{kw} {scoring}")


def test_real_reads_golden(oracle):
    """4000 of the 20,000 real read/window pairs the reference ships, against outputs of the reference kernels."""
    inp = np.load(os.path.join(HERE, "golden", "real4000_inputs.npz"))
    gold = np.load(os.path.join(HERE, "golden", "golden_real20k.npz"))
    n = 4000
    for name, kw, keys in (("local_start", dict(algo=po.LOCAL, start=po.WITH_START), K5), ("semi_start", dict(algo=po.SEMI_GLOBAL, start=po.WITH_START), K5),
                           ("global", dict(algo=po.GLOBAL), ["aln_score"]), ("banded_k64", dict(algo=po.BANDED, k_band=64), K3)):
        got = oracle.align(inp["packed_q"], inp["packed_t"], inp["q_lens"], inp["t_lens"], inp["q_offs"], inp["t_offs"], max_query_len=1024, **kw)
        want = {k: gold[f"{name}/{k}"][:n].astype(np.int32) for k in keys}
        assert_same(got, want, keys, f"real reads {name}")


def _one(oracle, q: bytes, t: bytes, **kw):
    b = synth.batch_from_sequences([q], [t])
    r = oracle.align(oracle.pack(b.q_bytes), oracle.pack(b.t_bytes), b.q_lens, b.t_lens, b.q_offs, b.t_offs, max_query_len=1024, **kw)
    return {k: int(v[0]) for k, v in r.items() if k not in ("cigar",)}


def test_quirk_q3_mirrored_target_start(oracle):
    # LOCAL WITH_START reports target start mirrored inside its 8-base word: 16*floor(ts/8) + 14 - ts
    motif = b"ACGTTGCAAGGCTTAACCGGTTAGCATGCAAT"
    for ts in (0, 1, 3, 5, 7, 9, 12):
        t = b"T" * 0 + bytes([ord("C")] * ts) + motif + b"G" * 11
        # make the flanks unable to extend the alignment
        t = (b"N" * ts) + motif + b"N" * 11
        r = _one(oracle, motif, t, algo=po.LOCAL, start=po.WITH_START)
        assert r["aln_score"] == len(motif)
        assert r["query_batch_start"] == 0
        assert r["target_batch_start"] == 16 * (ts // 8) + 14 - ts
        tb = _one(oracle, motif, t, algo=po.LOCAL, start=po.WITH_TB)
        assert tb["target_batch_start"] == ts and tb["query_batch_start"] == 0      # the traceback path reports the true start


def test_quirk_q4_global_top_boundary(oracle):
    # two extra leading query bases + 31 matches score 24 (a textbook Gotoh would give 23)
    core = b"ACGTTGCAAGGCTTAACCGGTTAGCATGCAA"
    r = _one(oracle, b"GG" + core, core, algo=po.GLOBAL)
    assert r["aln_score"] == 24


def test_quirk_q8_semiglobal_default_coordinates(oracle):
    # HEAD=TAIL=TARGET reports query_batch_end = ref_len and query_batch_start = read_len-1-ref_len
    rng = np.random.default_rng(3)
    t = bytes(synth.ACGT[rng.integers(0, 4, size=200)])
    q = t[30:180]
    r = _one(oracle, q, t, algo=po.SEMI_GLOBAL, start=po.WITH_START)
    assert r["aln_score"] == 150 and r["query_batch_end"] == 200 and r["query_batch_start"] == 150 - 1 - 200
    assert r["target_batch_end"] == 179 and r["target_batch_start"] == 30


def test_revcomp_gpu_semantics(oracle):
    # reverse = nibble reversal of the PADDED word array, complement = A<->T, C<->G
    b = synth.batch_from_sequences([b"ACGTTGCAAC"], [b"AAAAAAAAACCCCCCCCCGT"])
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    rq, rt = oracle.revcomp(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, np.array([3], np.uint8), np.array([1], np.uint8))
    # query padded: ACGTTGCA ACNNNNNN -> reversed NNNNNNCA ACGTTGCA -> complemented NNNNNNGT TGCAACGT
    assert rq.tolist() == [int("EEEEEE74", 16), int("47311374", 16)]
    want_t = oracle.pack(np.frombuffer((b"AAAAAAAAACCCCCCCCCGTNNNN")[::-1], dtype=np.uint8))
    assert np.array_equal(rt, want_t)


def test_banded_tb_extension_with_full_band_equals_local_tb(oracle):
    """oracle-only pin of the BANDED + WITH_TB extension: with a band wider than the matrix it IS the LOCAL WITH_TB restatement
    (which tests above pin to the reference's own kernels)"""
    from genomics_gpu_b200 import synth
    b = synth.random_ragged_pairs(300, 9, 200, 9, 300, seed=321, n_frac=0.01)
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    loc = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.LOCAL, start=po.WITH_TB)
    ban = oracle.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.BANDED, start=po.WITH_TB, k_band=8192)
    for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start", "target_batch_start", "n_cigar_ops"):
        assert np.array_equal(loc[k], ban[k]), k
    assert np.array_equal(loc["cigar"], ban["cigar"])
