"""Per-pair kernel choice inside ONE gasal_aln_async submission ("batch-granular dispatch"): a plain GASAL2 caller
may mix short and long pairs in a batch; the short ones must stay on the s16x2 inter-task fast path while the long ones run on the
intra-task wavefront kernel, and every result must still equal the oracle's in the caller's pair order."""
import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
from oracle import pyoracle as po

from helpers import K3, K5, assert_same, oracle_align

pytestmark = pytest.mark.gpu


def mixed_batch(seed, n_short=1500, n_long=40, n_frac=0.003):
    rng = np.random.default_rng(seed)
    short = synth.random_ragged_pairs(n_short, 9, 330, 9, 700, seed=seed, n_frac=n_frac)
    qs = [short.query(i) for i in range(short.n)]; ts = [short.target(i) for i in range(short.n)]
    for _ in range(n_long):
        L = int(rng.integers(400, 3000))
        t, q = synth.nw_style_pair(rng, L)
        pos = int(rng.integers(0, len(qs) + 1))
        qs.insert(pos, q); ts.insert(pos, t)
    return synth.batch_from_sequences(qs, ts)


def run(batch, **kw):
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
        res = al.align(batch)
        res["_kernel"] = al.kernel_name(0)
        res2 = al.align(batch)                       # the storage (and its per-class scratch) is reusable
        for k in res2:
            assert np.array_equal(res2[k], res[k]), k
    return res


def test_mixed_local_with_start(oracle):
    b = mixed_batch(6100)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START, max_query_len=4096)
    got = run(b, algo=capi.LOCAL, start=capi.WITH_START)
    assert got["_kernel"].startswith("mixed(s16x2+wave"), got["_kernel"]
    assert_same(got, want, K5, "mixed local WITH_START")


@pytest.mark.parametrize("head", [capi.TARGET, capi.BOTH])
def test_mixed_semiglobal_with_start(oracle, head):
    b = mixed_batch(6200 + head)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, head=head, start=po.WITH_START, max_query_len=4096)
    got = run(b, algo=capi.SEMI_GLOBAL, head=head, start=capi.WITH_START)
    assert got["_kernel"].startswith("mixed(s16x2+wave"), got["_kernel"]
    assert_same(got, want, K5, f"mixed semi-global head={head} WITH_START")


def test_mixed_global(oracle):
    b = mixed_batch(6300)
    want = oracle_align(oracle, b, algo=po.GLOBAL, max_query_len=4096)
    got = run(b, algo=capi.GLOBAL)
    assert got["_kernel"].startswith("mixed(s16x2+wave"), got["_kernel"]
    assert_same(got, want, ["aln_score"], "mixed global")


def test_mixed_banded(oracle):
    """BANDED score/end: the short class on the s16x2 banded sweep, the long class on the int32 banded kernel"""
    b = mixed_batch(6350, n_short=1200, n_long=25)
    want = oracle_align(oracle, b, algo=po.BANDED, k_band=48, max_query_len=4096)
    got = run(b, algo=capi.BANDED, k_band=48)
    assert got["_kernel"] == "mixed(s16x2+int32)", got["_kernel"]
    assert_same(got, want, K3, "mixed banded")


def test_one_long_read_does_not_take_a_short_read_batch_off_the_fast_path(oracle):
    """20,000 short pairs + ONE 5 kb pair: the short pairs stay on the s16x2 kernel"""
    rng = np.random.default_rng(7)
    short = synth.config2(20_000, seed=61)
    t, q = synth.nw_style_pair(rng, 5000)
    b = synth.batch_from_sequences([short.query(i) for i in range(short.n)] + [q], [short.target(i) for i in range(short.n)] + [t])
    want = oracle_align(oracle, b, algo=po.LOCAL, max_query_len=8192)
    got = run(b, algo=capi.LOCAL)
    assert got["_kernel"].startswith("mixed(s16x2+"), got["_kernel"]
    assert_same(got, want, K3, "20000 short + 1 long")


def test_modes_without_a_mixed_path_still_work(oracle):
    """second best / traceback on a mixed batch: the whole batch runs on the general kernel as before"""
    b = mixed_batch(6400, n_short=300, n_long=6)
    want = oracle_align(oracle, b, algo=po.LOCAL, second_best=True, max_query_len=4096)
    got = run(b, algo=capi.LOCAL, second_best=True)
    assert got["_kernel"] == "int32"
    assert_same(got, want, K3 + ["aln_score2", "query_batch_end2", "target_batch_end2"], "mixed batch, second best")
