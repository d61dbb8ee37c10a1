"""GPU parity of the nvbio-style BANDED scoring and TRACEBACK batches (include/nvbio_b200.h: nvbx_stream_enact_banded,
nvbx_stream_traceback, nvbx_stream_alignments) against oracle/nvbio_compat.c -- pinned to the reference's host build of
Non-CDP/NvB/nvbio/alignment and to its test program's known answers by tests/test_nvbio_tb_oracle.py -- and against the committed
golden vectors of that build.  Every call goes through the C ABI."""
import os

import numpy as np
import pytest

from genomics_gpu_b200 import nvb, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nvbio_tb_golden.npz")
TYPES = (nvb.GLOBAL, nvb.LOCAL, nvb.SEMI_GLOBAL)
ALIGNERS = [(nvb.GOTOH, (2, -1, -2, -1)), (nvb.GOTOH, (1, -3, -5, -2)), (nvb.GOTOH, (2, -1, -1, -1)), (nvb.SMITH_WATERMAN, (1, -1, -1, -2)),
            (nvb.SMITH_WATERMAN, (2, -1, -3, -1)), (nvb.EDIT_DISTANCE, (0, 0, 0, 0))]


@pytest.fixture(scope="module")
def orc():
    return po.NvbioOracle()


def jobs_of(pats, poffs, texts, toffs):
    return [(pats[poffs[k]:poffs[k + 1]], texts[toffs[k]:toffs[k + 1]]) for k in range(len(poffs) - 1)]


def rows_and_ops(alns, ops):
    rows = np.stack([alns[f].astype(np.int64) for f in ("score", "source_text", "source_pattern", "sink_text", "sink_pattern", "n_ops", "clip_end", "clip_begin")], axis=1)
    flat = np.concatenate([ops[k, :int(alns["n_ops"][k])] for k in range(len(alns))]) if len(alns) else np.zeros(0, np.uint8)
    return rows, flat


def want_rows(res):
    rows = np.array([[r["score"], *r["source"], *r["sink"], len(r["ops"]), r["clip_end"], r["clip_begin"]] for r in res], np.int64)
    return rows, np.concatenate([r["ops"] for r in res])


def test_golden_vectors_through_the_c_abi():
    g = np.load(GOLD)
    full = nvb.Batch(g["full_patterns"], g["full_poffs"], g["full_texts"], g["full_toffs"])
    band = nvb.Batch(g["band_patterns"], g["band_poffs"], g["band_texts"], g["band_toffs"])
    with nvb.Stream(full) as sf, nvb.Stream(band) as sb:
        for ci, cfg in enumerate(g["configs"]):
            a = [int(x) for x in cfg]
            for type_ in TYPES:
                al = nvb.aligner(a[0], type_, *a[1:])
                sf.traceback(al)
                rows, flat = rows_and_ops(*sf.alignments())
                assert np.array_equal(rows, g[f"full_{ci}_{type_}"]), (ci, type_, "full traceback")
                assert np.array_equal(flat, g[f"full_ops_{ci}_{type_}"]), (ci, type_, "full ops")
                for b in nvb.BAND_LENS:
                    sb.enact_banded(al, b, want_sinks=True)
                    s, si, sj = sb.results(True)
                    want = g[f"bscore_{ci}_{type_}_{b}"]
                    assert np.array_equal(s, want[:, 0]) and np.array_equal(si, want[:, 1]) and np.array_equal(sj, want[:, 2]), (ci, type_, b, "banded score")
                for b in (7, 15, 31):
                    sb.traceback(al, b)
                    rows, flat = rows_and_ops(*sb.alignments())
                    assert np.array_equal(rows, g[f"btb_{ci}_{type_}_{b}"]), (ci, type_, b, "banded traceback")
                    assert np.array_equal(flat, g[f"btb_ops_{ci}_{type_}_{b}"]), (ci, type_, b, "banded ops")


def test_known_answers_of_the_reference_test_program():
    """NvB/nvbio-test/alignment_test.cu:760-825 through the C ABI"""
    dna = lambda s: np.array(["ACGT".index(c) for c in s], dtype=np.uint8)
    p, t = dna("ACAACTA"), dna("AAACACCCTAACACACTAAA")
    with nvb.Stream(nvb.Batch(*nvb.concat([p]), *nvb.concat([t]))) as st:
        for algo in (nvb.SMITH_WATERMAN, nvb.GOTOH):
            for type_, want in ((nvb.GLOBAL, "1M2D3M1D3M10D"), (nvb.LOCAL, "4M1D3M"), (nvb.SEMI_GLOBAL, "4M1D3M")):
                st.traceback(nvb.aligner(algo, type_, 2, -1, -1, -1))
                alns, ops = st.alignments()
                assert po.nvb_rle(ops[0, :alns["n_ops"][0]]) == want, (algo, type_)
        st.traceback(nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -1, -1), 7)
        alns, ops = st.alignments()
        assert po.nvb_rle(ops[0, :alns["n_ops"][0]]) == "4M1D3M"
    p2 = dna("TTATGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTAT")
    t2 = dna("ATCGGATTCTTTCTTACTTGTAGGTGGTCTGGTTTTTGCCTTTTAAGCTTCTGCAAAAAACAACAACAAACTTGTGGTATTACACTGACTCTACAGATCAATTTGGGGACAACTTCCATGTGTTCCACCACCAATACTGAATCTTTCAATCGACTGACGTGGTAT" + "CTCTCTCTCCATCTAT")
    with nvb.Stream(nvb.Batch(*nvb.concat([p2]), *nvb.concat([t2]))) as st:
        st.traceback(nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 0, -5, -8, -3), 31)
        alns, ops = st.alignments()
        assert int(alns["score"][0]) == -11 and po.nvb_rle(ops[0, :alns["n_ops"][0]]) == "147M2D3M"


@pytest.mark.parametrize("algo,scheme", ALIGNERS)
def test_full_traceback_ragged_jobs_vs_oracle(orc, algo, scheme):
    pats, poffs, texts, toffs = synth.nvb_jobs(1500, 1, 170, 1, 260, seed=61 + algo)
    jobs = jobs_of(pats, poffs, texts, toffs)
    with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
        for type_ in TYPES:
            st.traceback(nvb.aligner(algo, type_, *scheme))
            assert st.kernel() == "int32+tb"
            rows, flat = rows_and_ops(*st.alignments())
            wr, wf = want_rows([orc.traceback(algo, type_, scheme, p, t) for p, t in jobs])
            bad = np.nonzero((rows != wr).any(axis=1))[0]
            assert bad.size == 0, (type_, bad[:5], rows[bad[:3]], wr[bad[:3]])
            assert np.array_equal(flat, wf), (type_, "ops")


@pytest.mark.parametrize("band", nvb.BAND_LENS)
def test_banded_scores_and_tracebacks_vs_oracle(orc, band):
    pats, poffs, texts, toffs = synth.nvb_banded_jobs(1200, 1, 200, band, seed=70 + band)
    jobs = jobs_of(pats, poffs, texts, toffs)
    with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
        for algo, scheme in ALIGNERS:
            for type_ in TYPES:
                al = nvb.aligner(algo, type_, *scheme)
                st.enact_banded(al, band, want_sinks=True)
                assert st.kernel() == "banded-int32"
                s, si, sj = st.results(True)
                want = np.array([orc.banded_score(algo, type_, band, scheme, p, t) for p, t in jobs], np.int64)
                assert np.array_equal(s, want[:, 0]) and np.array_equal(si, want[:, 1]) and np.array_equal(sj, want[:, 2]), (algo, scheme, type_)
                st.traceback(al, band)
                rows, flat = rows_and_ops(*st.alignments())
                wr, wf = want_rows([orc.traceback(algo, type_, scheme, p, t, band=band) for p, t in jobs])
                bad = np.nonzero((rows != wr).any(axis=1))[0]
                assert bad.size == 0, (algo, scheme, type_, bad[:5], rows[bad[:3]], wr[bad[:3]])
                assert np.array_equal(flat, wf), (algo, scheme, type_, "ops")


def test_banded_short_texts_indexed_pairs_and_ascii(orc):
    """texts shorter than pattern + band - 1 and shorter than the pattern (score defined, nothing reported), pair index arrays,
    ASCII symbols on the unpacked caches (band 15)"""
    rng = np.random.default_rng(5)
    pats, texts = [], []
    for _ in range(400):
        m = int(rng.integers(1, 90)); n = max(30, m + int(rng.integers(-4, 20)))
        t = rng.integers(0, 4, n).astype(np.uint8)
        p = t[:m].copy() if (n >= m and rng.random() < 0.6) else rng.integers(0, 4, m).astype(np.uint8)
        pats.append(p); texts.append(t)
    pp = rng.integers(0, 400, 900).astype(np.uint32); pt = rng.integers(0, 400, 900).astype(np.uint32)
    P, po_, T, to_ = *nvb.concat(pats), *nvb.concat(texts)
    with nvb.Stream(nvb.Batch(P, po_, T, to_, pair_pattern=pp, pair_text=pt)) as st:
        for band in (7, 31):
            for algo, scheme in ((nvb.GOTOH, (2, -1, -2, -1)), (nvb.EDIT_DISTANCE, (0, 0, 0, 0))):
                for type_ in TYPES:
                    st.enact_banded(nvb.aligner(algo, type_, *scheme), band, want_sinks=True)
                    s, si, sj = st.results(True)
                    want = np.array([orc.banded_score(algo, type_, band, scheme, pats[a], texts[b]) for a, b in zip(pp, pt)], np.int64)
                    assert np.array_equal(s, want[:, 0]) and np.array_equal(si, want[:, 1]) and np.array_equal(sj, want[:, 2]), (band, algo, type_)
    asc = np.frombuffer(b"ACGT", dtype=np.uint8)
    pats_a, poffs, texts_a, toffs = synth.nvb_banded_jobs(300, 5, 120, 15, seed=3)
    with nvb.Stream(nvb.Batch(asc[pats_a], poffs, asc[texts_a], toffs, alphabet=nvb.ASCII)) as st:
        st.traceback(nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1), 15)
        rows, flat = rows_and_ops(*st.alignments())
        wr, wf = want_rows([orc.traceback(po.NVB_GOTOH, po.NVB_LOCAL, (2, -1, -2, -1), p, t, band=15) for p, t in jobs_of(pats_a, poffs, texts_a, toffs)])
        assert np.array_equal(rows, wr) and np.array_equal(flat, wf)


def test_traceback_in_several_chunks(orc, monkeypatch):
    """a flow-word budget that holds only a few dozen jobs (NVBX_TB_BUDGET_WORDS): the chunk loop, partial last chunk, long strings"""
    pats, poffs, texts, toffs = synth.nvb_jobs(333, 40, 300, 200, 900, seed=81)
    jobs = jobs_of(pats, poffs, texts, toffs)
    monkeypatch.setenv("NVBX_TB_BUDGET_WORDS", str(38 * 900 * 64))            # ~64 jobs of the largest shape per chunk
    with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
        for type_ in TYPES:
            st.traceback(nvb.aligner(nvb.GOTOH, type_, 2, -1, -2, -1))
            rows, flat = rows_and_ops(*st.alignments())
            wr, wf = want_rows([orc.traceback(po.NVB_GOTOH, type_, (2, -1, -2, -1), p, t) for p, t in jobs])
            assert np.array_equal(rows, wr) and np.array_equal(flat, wf), type_
    pats, poffs, texts, toffs = synth.nvb_banded_jobs(333, 40, 300, 31, seed=82)
    jobs = jobs_of(pats, poffs, texts, toffs)
    monkeypatch.setenv("NVBX_TB_BUDGET_WORDS", str(300 * 4 * 40))
    with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
        st.traceback(nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1), 31)
        rows, flat = rows_and_ops(*st.alignments())
        wr, wf = want_rows([orc.traceback(po.NVB_GOTOH, po.NVB_LOCAL, (2, -1, -2, -1), p, t, band=31) for p, t in jobs])
        assert np.array_equal(rows, wr) and np.array_equal(flat, wf)


def test_traceback_reuse_and_error_paths(orc):
    """a second traceback on the same stream (buffers reused), alignments() before any traceback, unknown band lengths"""
    pats, poffs, texts, toffs = synth.nvb_jobs(300, 10, 60, 20, 90, seed=4)
    with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
        with pytest.raises(RuntimeError, match="no traceback"):
            st.alignments()
        with pytest.raises(RuntimeError, match="band_len"):
            st.enact_banded(nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1), 8)
        with pytest.raises(RuntimeError, match="band_len"):
            st.traceback(nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1), 4)
        al = nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1)
        st.traceback(al); first = rows_and_ops(*st.alignments())
        st.enact(al); st.results()                                       # a scoring pass in between leaves no traceback to read
        with pytest.raises(RuntimeError, match="no traceback"):
            st.alignments()
        st.traceback(al); second = rows_and_ops(*st.alignments())
        assert np.array_equal(first[0], second[0]) and np.array_equal(first[1], second[1])
        jobs = jobs_of(pats, poffs, texts, toffs)
        wr, wf = want_rows([orc.traceback(po.NVB_GOTOH, po.NVB_SEMI_GLOBAL, (2, -1, -2, -1), p, t) for p, t in jobs])
        assert np.array_equal(first[0], wr) and np.array_equal(first[1], wf)


def test_cpp_front_end_traceback(orc, tmp_path):
    """aln::BatchedAlignmentTraceback of include/nvbio_b200_batched.hpp through tools/nvb_cpp_example.cpp (second argument: also trace back)"""
    import subprocess
    from test_nvbio_oracle import _write_jobs
    import __graft_entry__ as g
    exe = g.build_cpp_example()
    pats, poffs, texts, toffs = synth.nvb_jobs(201, 5, 90, 30, 160, seed=45)
    _write_jobs(tmp_path / "jobs.bin", pats, poffs, texts, toffs)
    r = subprocess.run([exe, str(tmp_path / "jobs.bin"), "tb"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln.split(" ") for ln in r.stdout.strip().split("\n") if ln.startswith("4 ")]
    assert len(lines) == 201
    for k, (p, t) in enumerate(jobs_of(pats, poffs, texts, toffs)):
        w = orc.traceback(po.NVB_GOTOH, po.NVB_LOCAL, (2, -1, -2, -1), p, t)
        assert [int(x) for x in lines[k][1:7]] == [w["score"], *w["source"], *w["sink"], len(w["ops"])], k
        assert lines[k][7] == "".join("MID"[o] for o in w["ops"]), k
