"""nvbio-style BANDED scoring and TRACEBACK (the rest of SURVEY.md 8f row f4): the restatement in oracle/nvbio_compat.c against
(1) the known answers of the reference's own test program (NvB/nvbio-test/alignment_test.cu:676-825), (2) the committed golden
vectors generated from the reference's host build (tests/golden/make_nvbio_tb_golden.py) and (3), where /root/reference was
available to build oracle/_ref/libnvbio_ref.so, the reference's host code path itself on random jobs."""
import os

import numpy as np
import pytest

from genomics_gpu_b200 import synth
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nvbio_tb_golden.npz")
TYPES = (po.NVB_GLOBAL, po.NVB_LOCAL, po.NVB_SEMI_GLOBAL)


@pytest.fixture(scope="module")
def orc():
    return po.NvbioOracle()


def dna(s):
    return np.array(["ACGT".index(c) for c in s], dtype=np.uint8)          # string_to_dna of the reference's tests


def row_of(r):
    return [r["score"], *r["source"], *r["sink"], len(r["ops"]), r["clip_end"], r["clip_begin"]]


def test_known_answers_of_the_reference_test_program(orc):
    """alignment_test.cu:760-825: the expected strings are run-length encodings of the ops in push order"""
    p, t = dna("ACAACTA"), dna("AAACACCCTAACACACTAAA")
    for algo in (po.NVB_SW, po.NVB_GOTOH):                                  # 2 / -1 / -1 / -1 in both schemes (:768-771, :780-783)
        sc = (2, -1, -1, -1)
        assert po.nvb_rle(orc.traceback(algo, po.NVB_GLOBAL, sc, p, t)["ops"]) == "1M2D3M1D3M10D"
        assert po.nvb_rle(orc.traceback(algo, po.NVB_LOCAL, sc, p, t)["ops"]) == "4M1D3M"
        assert po.nvb_rle(orc.traceback(algo, po.NVB_SEMI_GLOBAL, sc, p, t)["ops"]) == "4M1D3M"
    assert po.nvb_rle(orc.traceback(po.NVB_GOTOH, po.NVB_SEMI_GLOBAL, (2, -1, -1, -1), p, t, band=7)["ops"]) == "4M1D3M"      # :793
    # "real banded Gotoh problem" (:797-825): BAND_LEN 31, 0 / -5 / -8 / -3
    p2 = dna("TTATGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTAT")
    t2 = dna("ATCGGATTCTTTCTTACTTGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTATCTCTCTCTCCATCTAT")
    r = orc.traceback(po.NVB_GOTOH, po.NVB_SEMI_GLOBAL, (0, -5, -8, -3), p2, t2, band=31)
    assert po.nvb_rle(r["ops"]) == "147M2D3M" and r["score"] == -11
    assert orc.banded_score(po.NVB_GOTOH, po.NVB_SEMI_GLOBAL, 31, (0, -5, -8, -3), p2, t2)[0] == -11
    # banded edit distance, BAND_LEN 5, semi-global (:676-747)
    asc = lambda s: np.frombuffer(s.encode(), dtype=np.uint8)
    for pat, txt, want in (("GGGTGCTCAA", "AAAAGGGTGCTCAA", 0), ("GGGTAAGCTC", "AAAAGGGTGCTCAA", -2), ("AAGGGTGCTC", "AAAAGGGTGCAATC", -2),
                           ("AAAAGGGTGC", "AAAAGGGTGCTCAA", 0), ("AAAAGGGTG", "AAAAGGAAGTGCTC", -2), ("CACCGGGT", "AACAGGGTGCTC", -2)):
        assert orc.banded_score(po.NVB_ED, po.NVB_SEMI_GLOBAL, 5, (), asc(pat), asc(txt))[0] == want


def test_golden_vectors_from_the_reference_build(orc):
    g = np.load(GOLD)
    fp, fpo, ft, fto = g["full_patterns"], g["full_poffs"], g["full_texts"], g["full_toffs"]
    bp, bpo, bt, bto = g["band_patterns"], g["band_poffs"], g["band_texts"], g["band_toffs"]
    full_jobs = [(fp[fpo[k]:fpo[k + 1]], ft[fto[k]:fto[k + 1]]) for k in range(len(fpo) - 1)]
    band_jobs = [(bp[bpo[k]:bpo[k + 1]], bt[bto[k]:bto[k + 1]]) for k in range(len(bpo) - 1)]
    for ci, cfg in enumerate(g["configs"]):
        algo, scheme = int(cfg[0]), tuple(int(x) for x in cfg[1:])
        for type_ in TYPES:
            res = [orc.traceback(algo, type_, scheme, p, t) for p, t in full_jobs]
            assert np.array_equal(np.array([row_of(r) for r in res], np.int32), g[f"full_{ci}_{type_}"]), (algo, type_, "full traceback")
            assert np.array_equal(np.concatenate([r["ops"] for r in res]), g[f"full_ops_{ci}_{type_}"]), (algo, type_, "full ops")
            for band in (3, 5, 7, 15, 31):
                got = np.array([orc.banded_score(algo, type_, band, scheme, p, t) for p, t in band_jobs], np.int32)
                assert np.array_equal(got, g[f"bscore_{ci}_{type_}_{band}"]), (algo, type_, band, "banded score")
            for band in (7, 15, 31):
                res = [orc.traceback(algo, type_, scheme, p, t, band=band) for p, t in band_jobs]
                assert np.array_equal(np.array([row_of(r) for r in res], np.int32), g[f"btb_{ci}_{type_}_{band}"]), (algo, type_, band, "banded traceback")
                assert np.array_equal(np.concatenate([r["ops"] for r in res]), g[f"btb_ops_{ci}_{type_}_{band}"]), (algo, type_, band, "banded ops")


def _same(a, b):
    return row_of(a) == row_of(b) and np.array_equal(a["ops"], b["ops"])


@pytest.mark.parametrize("algo,scheme", [(po.NVB_GOTOH, (2, -1, -2, -1)), (po.NVB_GOTOH, (1, -3, -5, -2)), (po.NVB_GOTOH, (2, -1, -1, -1)),
                                         (po.NVB_SW, (1, -1, -1, -2)), (po.NVB_SW, (1, -2, -3, -1)), (po.NVB_ED, ())])
def test_restatement_equals_reference_host_path(orc, algo, scheme):
    if orc.ref is None or not hasattr(orc.ref, "nvbio_ref_traceback"):
        pytest.skip("oracle/_ref/libnvbio_ref.so not built (no /root/reference here)")
    pats, poffs, texts, toffs = synth.nvb_jobs(120, 1, 140, 1, 200, seed=31 + algo)
    for type_ in TYPES:
        for k in range(len(poffs) - 1):
            p, t = pats[poffs[k]:poffs[k + 1]], texts[toffs[k]:toffs[k + 1]]
            assert _same(orc.traceback(algo, type_, scheme, p, t), orc.traceback(algo, type_, scheme, p, t, ref=True)), (type_, k)
    for band in (3, 5, 7, 15, 31):
        pats, poffs, texts, toffs = synth.nvb_banded_jobs(60, 1, 130, band, seed=41 + band)
        for type_ in TYPES:
            for k in range(len(poffs) - 1):
                p, t = pats[poffs[k]:poffs[k + 1]], texts[toffs[k]:toffs[k + 1]]
                assert orc.banded_score(algo, type_, band, scheme, p, t) == orc.banded_score(algo, type_, band, scheme, p, t, ref=True), (band, type_, k)
                if band >= 7:
                    assert _same(orc.traceback(algo, type_, scheme, p, t, band=band), orc.traceback(algo, type_, scheme, p, t, band=band, ref=True)), (band, type_, k)


def test_banded_scores_outside_the_traceback_domain_equal_the_reference(orc):
    """texts SHORTER than pattern + band - 1 (and shorter than the pattern): scoring is defined there (255 sentinel past the text)"""
    if orc.ref is None or not hasattr(orc.ref, "nvbio_ref_banded_score"):
        pytest.skip("oracle/_ref/libnvbio_ref.so not built (no /root/reference here)")
    rng = np.random.default_rng(9)
    for _ in range(300):
        band = int(rng.choice([3, 5, 7, 15, 31]))
        m = int(rng.integers(1, 80)); n = max(band - 1, m + int(rng.integers(-3, band)))
        t = rng.integers(0, 4, n).astype(np.uint8); p = rng.integers(0, 4, m).astype(np.uint8)
        if n >= m and rng.random() < 0.7:
            a = int(rng.integers(0, n - m + 1)); p = t[a:a + m].copy()
        for algo, scheme in ((po.NVB_GOTOH, (2, -1, -2, -1)), (po.NVB_SW, (1, -1, -1, -2)), (po.NVB_ED, ())):
            for type_ in TYPES:
                assert orc.banded_score(algo, type_, band, scheme, p, t) == orc.banded_score(algo, type_, band, scheme, p, t, ref=True), (band, m, n, algo, type_)
