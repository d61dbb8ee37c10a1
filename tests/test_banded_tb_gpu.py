"""Banded alignment with CIGAR traceback (BASELINE.json configs[3]; north_star item 4).

The reference has NO banded traceback: gasal_banded_tiled_kernel returns score/end only.  The engine's
BANDED + WITH_TB mode is defined as the reference's LOCAL recurrence / direction nibbles / walker restricted to the reference's
tile band, so parity is pinned the way SURVEY 8c prescribes:
  (i)   score/end without traceback against the banded kernel restatement (tests/test_parity_gpu.py::test_banded);
  (ii)  CIGAR, starts, score, ends against the reference's UNBANDED LOCAL WITH_TB wherever the band does not cut the alignment;
  (iii) every CIGAR re-scores to aln_score;
and bit-exact against the oracle's restatement of the same definition for every pair (band-cut ones included)."""
import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
from oracle import pyoracle as po

from helpers import K5, assert_same, oracle_align, cigar_slices, cigar_rescore

pytestmark = pytest.mark.gpu


def run(batch, **kw):
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
        return al.align(batch)


@pytest.mark.parametrize("k_band", [16, 32, 64, 4096])
def test_banded_tb_matches_oracle(oracle, k_band):
    b = synth.random_ragged_pairs(400, 9, 220, 9, 330, seed=500 + k_band, n_frac=0.01)
    want = oracle_align(oracle, b, algo=po.BANDED, start=po.WITH_TB, k_band=k_band)
    got = run(b, algo=capi.BANDED, start=capi.WITH_TB, k_band=k_band)
    assert_same(got, want, K5 + ["n_cigar_ops"], f"banded tb k={k_band}")
    assert cigar_slices(got, b) == cigar_slices(want, b)


def test_banded_tb_config4_equals_reference_local_tb(oracle):
    """C4 shape: 250 bp reads against the read locus +- 24 bp, k_band = 64: the band never cuts the alignment, so the
    banded result must equal the reference's unbanded LOCAL WITH_TB (score, ends, starts, CIGAR)."""
    b = synth.config4(3000, seed=4)
    ref = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_TB)
    got = run(b, algo=capi.BANDED, start=capi.WITH_TB, k_band=64)
    assert_same(got, ref, K5 + ["n_cigar_ops"], "banded tb (k=64) vs unbanded local tb")
    gc, rc = cigar_slices(got, b), cigar_slices(ref, b)
    assert gc == rc
    scoring = (1, 4, 6, 1)
    for i, c in enumerate(gc):
        if got["aln_score"][i] > 0:
            assert cigar_rescore(c, scoring) == got["aln_score"][i], f"pair {i}"


def test_banded_tb_narrow_band_rescoring(oracle):
    """a band that does cut alignments: results still form a consistent alignment (CIGAR re-scores to the reported score,
    spans exactly start..end) and never beat the unbanded optimum"""
    rng = np.random.default_rng(77)
    qs, ts = [], []
    for _ in range(600):
        core = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(160, 260)))
        cut = int(rng.integers(40, core.size - 40)); gap = int(rng.integers(24, 64))
        ins = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=gap)
        qs.append(core.tobytes())                                              # the target carries a long insertion:
        ts.append(np.concatenate([core[:cut], ins, core[cut:]]).tobytes())    # the full alignment needs a gap wider than the band
    b = synth.batch_from_sequences(qs, ts)
    full = oracle_align(oracle, b, algo=po.LOCAL)
    got = run(b, algo=capi.BANDED, start=capi.WITH_TB, k_band=8)
    assert np.all(got["aln_score"] <= full["aln_score"])
    gc = cigar_slices(got, b)
    for i, c in enumerate(gc):
        if got["aln_score"][i] <= 0:
            continue
        assert cigar_rescore(c, (1, 4, 6, 1)) == got["aln_score"][i]
        q_span = sum(x >> 2 for x in c if (x & 3) != 2); t_span = sum(x >> 2 for x in c if (x & 3) != 3)
        # the walker stops ON the first aligned cell (get_tb.h:100-103,142-145): start..end inclusive
        assert got["query_batch_end"][i] - got["query_batch_start"][i] + 1 == q_span
        assert got["target_batch_end"][i] - got["target_batch_start"][i] + 1 == t_span


def test_banded_tb_runs_on_the_packed_kernel(oracle):
    """north_star (4): BANDED + WITH_TB on the s16x2 lanes (k_banded_tb_s16), incl. pairs with N (N-aware pass) and odd band geometry"""
    b = synth.with_n_bases(synth.config4(4096, seed=45), 0.1, seed=3)
    want = oracle_align(oracle, b, algo=po.BANDED, start=po.WITH_TB, k_band=64)
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    with Aligner(algo=capi.BANDED, start=capi.WITH_TB, k_band=64, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
        got = al.align(b)
        assert al.kernel_name(0) == "s16x2"
    assert_same(got, want, K5 + ["n_cigar_ops"], "banded tb s16x2, C4 shape with N")
    assert cigar_slices(got, b) == cigar_slices(want, b)
    for k_band, qmax, tmax in ((8, 40, 330), (24, 220, 60), (40, 330, 330), (4096, 100, 200)):
        b = synth.random_ragged_pairs(500, 9, qmax, 9, tmax, seed=600 + k_band, n_frac=0.01)
        want = oracle_align(oracle, b, algo=po.BANDED, start=po.WITH_TB, k_band=k_band)
        got = run(b, algo=capi.BANDED, start=capi.WITH_TB, k_band=k_band)
        assert_same(got, want, K5 + ["n_cigar_ops"], f"banded tb s16x2 k={k_band} q<={qmax} t<={tmax}")
        assert cigar_slices(got, b) == cigar_slices(want, b)
