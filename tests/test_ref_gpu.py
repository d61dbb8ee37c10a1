"""GPU tests against the REFERENCE's own CUDA build (oracle/_ref/ref_gpu_driver_mq*, built from the reference
sources for sm_100 by oracle/Makefile): three-way agreement reference-GPU == oracle == our engine."""
import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
from oracle import pyoracle as po

from helpers import K3, K5, K2ND, assert_same, oracle_align, cigar_slices

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not po.RefGpu.available(608), reason="reference GPU build not present in oracle/_ref")]


def ours(batch, **kw):
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
        return al.align(batch, q_op=kw.get("_q_op"), t_op=kw.get("_t_op"))


@pytest.fixture(scope="module")
def batch():
    # q <= 150, t <= 600 so that MAX_QUERY_LEN=608 also covers the semi-global start pass
    return synth.random_ragged_pairs(1500, 30, 150, 30, 600, seed=4242, n_frac=0.01)


def test_local_three_way(oracle, batch):
    ref = po.RefGpu(608).run(batch, algo=po.LOCAL, start=po.WITH_START, batch_size=400)
    orc = oracle_align(oracle, batch, algo=po.LOCAL, start=po.WITH_START)
    assert_same(orc, ref, K5, "oracle vs reference GPU (local with start)")
    got = ours(batch, algo=capi.LOCAL, start=capi.WITH_START)
    assert_same(got, ref, K5, "engine vs reference GPU (local with start)")


def test_semiglobal_three_way(oracle, batch):
    ref = po.RefGpu(608).run(batch, algo=po.SEMI_GLOBAL, start=po.WITH_START, batch_size=400)
    orc = oracle_align(oracle, batch, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=608)
    assert_same(orc, ref, K5, "oracle vs reference GPU (semi-global with start)")
    got = ours(batch, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
    assert_same(got, ref, K5, "engine vs reference GPU (semi-global with start)")


def test_global_and_banded_three_way(oracle, batch):
    ref = po.RefGpu(608).run(batch, algo=po.GLOBAL, batch_size=400)
    assert_same(oracle_align(oracle, batch, algo=po.GLOBAL), ref, ["aln_score"], "oracle vs reference GPU (global)")
    assert_same(ours(batch, algo=capi.GLOBAL), ref, ["aln_score"], "engine vs reference GPU (global)")
    ref = po.RefGpu(608).run(batch, algo=po.BANDED, k_band=48, batch_size=400)
    assert_same(oracle_align(oracle, batch, algo=po.BANDED, k_band=48), ref, K3, "oracle vs reference GPU (banded)")
    assert_same(ours(batch, algo=capi.BANDED, k_band=48), ref, K3, "engine vs reference GPU (banded)")


def test_local_traceback_three_way(oracle):
    b = synth.random_ragged_pairs(600, 30, 150, 30, 300, seed=77, n_frac=0.0)
    ref = po.RefGpu(608).run(b, algo=po.LOCAL, start=po.WITH_TB, batch_size=600)      # one batch: CIGAR slots are per batch
    orc = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_TB)
    assert_same(orc, ref, K5 + ["n_cigar_ops"], "oracle vs reference GPU (local tb)")
    got = ours(b, algo=capi.LOCAL, start=capi.WITH_TB)
    assert_same(got, ref, K5 + ["n_cigar_ops"], "engine vs reference GPU (local tb)")
    assert cigar_slices(got, b) == cigar_slices(ref, b) == cigar_slices(orc, b)


def test_reverse_complement_three_way(oracle):
    """gasal_reversecomplement_kernel shifts words by 32 (undefined in C).  nvcc 12.9 folds the undefined
    operand to 0 for sm_100 (profiles/reference_revcomp_kernel_sm100_nvcc12.9.sass); revcomp mode 1 reproduces exactly
    that, mode 0 (default) is the clean padded-array reversal.  Complement-only ops are well defined and must agree in
    both modes."""
    b = synth.random_ragged_pairs(800, 9, 150, 9, 400, seed=99)
    rng = np.random.default_rng(7)
    q_op = rng.integers(0, 4, size=b.n).astype(np.uint8); t_op = rng.integers(0, 4, size=b.n).astype(np.uint8)
    ref = po.RefGpu(608).run(b, algo=po.LOCAL, start=po.WITH_START, q_op=q_op, t_op=t_op, batch_size=800)
    pq, pt = oracle.pack(b.q_bytes), oracle.pack(b.t_bytes)
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    lib = capi.load()
    for mode in (1, 0):
        rq, rt = oracle.revcomp(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, q_op, t_op, mode=mode)
        orc = oracle.align(rq, rt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.LOCAL, start=po.WITH_START)
        lib.gasalx_set_revcomp_mode(mode)
        try:
            with Aligner(algo=capi.LOCAL, start=capi.WITH_START, reverse_complement=True, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
                got = al.align(b, q_op=q_op, t_op=t_op)
        finally:
            lib.gasalx_set_revcomp_mode(0)
        assert_same(got, orc, K5, f"engine vs oracle, revcomp mode {mode}")
        # mode 1: every op; mode 0: only the pairs without a reverse bit are comparable with this reference build
        ok = np.ones(b.n, bool) if mode == 1 else ((q_op & 1) == 0) & ((t_op & 1) == 0)
        assert ok.sum() > 100
        for k in K5:
            assert np.array_equal(got[k][ok], ref[k][ok]), f"engine vs reference GPU, revcomp mode {mode}: {k}"
