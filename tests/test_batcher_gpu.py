"""GPU tests of the library-side batcher (gasalx_align_all): a mixed-length job, several worker threads (mapped onto
the one test GPU), results must come back bit-exact in the caller's order."""
import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import align_all
from oracle import pyoracle as po

from helpers import K3, K5, assert_same, oracle_align, cigar_slices

pytestmark = pytest.mark.gpu


def mixed_job(seed=31, n=1500):
    """short reads and a tail of long pairs: exercises the fast path, the general kernels and the bin boundary"""
    a = synth.random_ragged_pairs(n, 30, 250, 30, 600, seed=seed, n_frac=0.005)
    b = synth.random_ragged_pairs(60, 400, 1500, 400, 2000, seed=seed + 1, related=0.9, sub=0.08, indel=0.04, repeat_frac=0.0)
    rng = np.random.default_rng(seed)
    qs = [a.query(i) for i in range(a.n)] + [b.query(i) for i in range(b.n)]
    ts = [a.target(i) for i in range(a.n)] + [b.target(i) for i in range(b.n)]
    perm = rng.permutation(len(qs))
    return synth.batch_from_sequences([qs[i] for i in perm], [ts[i] for i in perm])


@pytest.mark.parametrize("workers", [1, 3])
def test_align_all_local_with_start(oracle, workers):
    job = mixed_job()
    want = oracle_align(oracle, job, algo=po.LOCAL, start=po.WITH_START)
    got = align_all(job, algo=capi.LOCAL, start=capi.WITH_START, n_gpus=workers, gpu_ids=[0] * workers, max_batch_pairs=300, streams_per_gpu=2)
    assert_same(got, want, K5, f"align_all local/start, {workers} workers")


def test_align_all_semiglobal_global_banded(oracle):
    job = mixed_job(seed=77, n=800)
    want = oracle_align(oracle, job, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=4096)
    got = align_all(job, algo=capi.SEMI_GLOBAL, start=capi.WITH_START, n_gpus=2, gpu_ids=[0, 0], max_batch_pairs=256)
    assert_same(got, want, K5, "align_all semi-global")
    want = oracle_align(oracle, job, algo=po.GLOBAL)
    got = align_all(job, algo=capi.GLOBAL, n_gpus=2, gpu_ids=[0, 0], max_batch_pairs=256)
    assert_same(got, want, ["aln_score"], "align_all global")
    want = oracle_align(oracle, job, algo=po.BANDED, k_band=64)
    got = align_all(job, algo=capi.BANDED, k_band=64, n_gpus=2, gpu_ids=[0, 0], max_batch_pairs=256)
    assert_same(got, want, K3, "align_all banded")


def test_align_all_traceback(oracle):
    job = synth.random_ragged_pairs(700, 30, 200, 30, 300, seed=8)
    want = oracle_align(oracle, job, algo=po.LOCAL, start=po.WITH_TB)
    got = align_all(job, algo=capi.LOCAL, start=capi.WITH_TB, n_gpus=2, gpu_ids=[0, 0], max_batch_pairs=200)
    assert_same(got, want, K5 + ["n_cigar_ops"], "align_all local tb")
    assert cigar_slices(got, job) == cigar_slices(want, job)


def test_align_all_input_order_mode(oracle):
    job = mixed_job(seed=5, n=400)
    want = oracle_align(oracle, job, algo=po.LOCAL)
    got = align_all(job, algo=capi.LOCAL, n_gpus=1, max_batch_pairs=128, keep_input_order=True)
    assert_same(got, want, K3, "align_all keep_input_order")


def test_fasta_files_to_cigar_strings_end_to_end(oracle, tmp_path):
    """the reference caller's whole flow without its host code: FASTA files -> native reader -> batcher -> CIGAR strings
    (test_prog.cpp:79-137, 203-437); header marks drive reverse / complement (modes agree for complement-only ops)"""
    from genomics_gpu_b200.aligner import read_fasta_pairs, cigar_string
    b = synth.random_ragged_pairs(500, 20, 200, 20, 300, seed=12)
    marks = [">", "/"] * 250                                       # forward / complement (well defined in the reference)
    for name, seqs, mk in (("q.fa", [b.query(i) for i in range(b.n)], marks), ("t.fa", [b.target(i) for i in range(b.n)], [">"] * b.n)):
        with open(tmp_path / name, "w") as f:
            for i, s in enumerate(seqs):
                f.write(f"{mk[i]}r{i}\n{s.decode()}\n")
    job, q_op, t_op, qn, tn = read_fasta_pairs(str(tmp_path / "q.fa"), str(tmp_path / "t.fa"))
    assert job.n == b.n and qn[3] == "r3"
    pq, pt = oracle.pack(job.q_bytes), oracle.pack(job.t_bytes)
    q32, t32 = job.q_offs.astype(np.uint32), job.t_offs.astype(np.uint32)
    rq, rt = oracle.revcomp(pq, pt, job.q_lens, job.t_lens, q32, t32, q_op, t_op)
    want = oracle.align(rq, rt, job.q_lens, job.t_lens, q32, t32, algo=po.LOCAL, start=po.WITH_TB)
    got = align_all(job, algo=capi.LOCAL, start=capi.WITH_TB, n_gpus=2, gpu_ids=[0, 0], max_batch_pairs=128, q_op=q_op, t_op=t_op)
    assert_same(got, want, K5 + ["n_cigar_ops"], "fasta -> align_all -> tb")
    for i in range(job.n):
        o, n = int(job.q_offs[i]), int(got["n_cigar_ops"][i])
        s = cigar_string(got["cigar"][o:o + n], n)
        assert s == cigar_string(want["cigar"][o:o + n], n)
        if got["aln_score"][i] > 0:
            assert sum(int(x) for x in __import__("re").findall(r"(\d+)[MXI]", s)) == got["query_batch_end"][i] - got["query_batch_start"][i] + 1
