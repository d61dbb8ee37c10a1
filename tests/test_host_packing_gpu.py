"""Host-side 4-bit packing inside gasal_host_batch_fill / _add / _addbase (default on, gasalx_set_host_packing): the caller hands over
ASCII exactly as with the reference (Non-CDP/GASAL2/src/host_batch.cpp:79-223), the pinned page and the H2D copy carry 4-bit words.
Both host layouts must give the oracle's results through every fill path, including page growth."""
import ctypes as C

import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
from oracle import pyoracle as po

from helpers import K5, assert_same, oracle_align

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 0], ids=["host_packed", "ascii_pages"])
def host_packing(request):
    lib = capi.load()
    lib.gasalx_set_host_packing(request.param)
    yield request.param
    lib.gasalx_set_host_packing(1)


def test_bulk_add_and_one_by_one_fill(oracle, host_packing):
    b = synth.random_ragged_pairs(700, 9, 220, 9, 330, seed=4100, n_frac=0.01)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START)
    mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
    lib = capi.load()
    with Aligner(algo=capi.LOCAL, start=capi.WITH_START, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=2) as al:
        assert lib.gasalx_storage_host_packed(C.byref(al.vec.a[0])) == host_packing
        got = al.align(b, i=0)                                         # two bulk gasal_host_batch_add calls
        assert_same(got, want, K5, "bulk add")
        al.fill_one_by_one(1, [b.query(i) for i in range(b.n)], [b.target(i) for i in range(b.n)])     # gasal_host_batch_fill per sequence
        qb, tb = al._last_bytes
        assert (qb, tb) == (b.q_bytes.size, b.t_bytes.size)
        al.launch(1, qb, tb, b.n); al.wait(1)
        assert_same(al.results(1, b.n), want, K5, "fill one by one")


def test_pages_grow_beyond_their_initial_size(oracle, host_packing):
    """storages sized for 50 x 64 bases receive 900 ragged pairs one sequence at a time: the host pages grow (linked 2x pages for
    ASCII, reallocation for packed pages), the device buffers regrow, results stay exact; then the storage is reused"""
    b = synth.random_ragged_pairs(900, 9, 220, 9, 330, seed=4200)
    want = oracle_align(oracle, b, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=1024)
    with Aligner(algo=capi.SEMI_GLOBAL, start=capi.WITH_START, max_query_len=64, max_target_len=64, max_n_alns=50, n_streams=1) as al:
        for rep in range(2):
            al.vec.a[0].current_n_alns = 0
            al.fill_one_by_one(0, [b.query(i) for i in range(b.n)], [b.target(i) for i in range(b.n)])
            qb, tb = al._last_bytes
            al.launch(0, qb, tb, b.n); al.wait(0)
            assert_same(al.results(0, b.n), want, K5, f"grown pages, pass {rep}")


def test_addbase_and_unaligned_add(oracle, host_packing):
    """gasal_host_batch_addbase / _add at arbitrary byte positions (host_batch.cpp:156-223): base-by-base assembly of a batch"""
    b = synth.random_ragged_pairs(60, 9, 90, 9, 120, seed=4300)
    want = oracle_align(oracle, b, algo=po.LOCAL)
    lib = capi.load()
    with Aligner(algo=capi.LOCAL, max_query_len=96, max_target_len=128, max_n_alns=b.n, n_streams=1) as al:
        s = al.vec.a[0]
        for src, data in ((capi.QUERY, b.q_bytes), (capi.TARGET, b.t_bytes)):
            idx = 0
            raw = data.tobytes()
            while idx < len(raw):                                        # alternate single bases and odd-sized runs
                if idx % 5 == 0:
                    idx = lib.gasal_host_batch_addbase(C.byref(s), idx, raw[idx:idx + 1], src)
                else:
                    n = min(7, len(raw) - idx)
                    idx = lib.gasal_host_batch_add(C.byref(s), idx, raw[idx:idx + n], n, src)
            assert idx == len(raw)
        n = b.n
        C.memmove(s.host_query_batch_offsets, b.q_offs.ctypes.data, 4 * n); C.memmove(s.host_target_batch_offsets, b.t_offs.ctypes.data, 4 * n)
        C.memmove(s.host_query_batch_lens, b.q_lens.ctypes.data, 4 * n); C.memmove(s.host_target_batch_lens, b.t_lens.ctypes.data, 4 * n)
        s.current_n_alns = n
        al.launch(0, b.q_bytes.size, b.t_bytes.size, n); al.wait(0)
        assert_same(al.results(0, n), want, ["aln_score", "query_batch_end", "target_batch_end"], "addbase / unaligned add")


def test_deprecated_gpu_mem_alloc_free_link_and_run():
    """gasal_gpu_mem_alloc / gasal_gpu_mem_free (reference src/ctors.h:9,11, deprecated there but declared): allocate and free the
    device side of a zeroed storage"""
    lib = capi.load()
    p = capi.Params(); lib.gasal_params_default(C.byref(p)); p.algo = capi.LOCAL
    v = lib.gasal_init_gpu_storage_v(1)
    s = v.a[0]
    s.device_cpy = lib.gasal_res_new_device_cpy(1000, C.byref(p))
    lib.gasal_gpu_mem_alloc(C.byref(s), 160 * 1000, 320 * 1000, 1000, C.byref(p))
    assert s.gpu_max_query_batch_bytes == 160000 and s.gpu_max_n_alns == 1000 and s.packed_query_batch and s.unpacked_target_batch
    lib.gasal_gpu_mem_free(C.byref(s), C.byref(p))
    assert not s.packed_query_batch and not s.query_batch_lens
    lib.gasal_destroy_gpu_storage_v(C.byref(v))


def test_clean_batches_skip_the_extra_passes(oracle):
    """host packing knows whether a batch holds any base outside A/C/G/T (pads excluded when gasal_host_batch_fill writes them): a clean
    batch runs ONE DP launch; the same batch with real N bases runs classification + the N-aware pass; both equal the oracle"""
    lib = capi.load()
    clean = synth.make_read_window_pairs(3000, 152, 304, sub=0.03, indel=0.003, seed=4400)        # lengths are multiples of 8: no pads at all
    dirty = synth.with_n_bases(clean, 0.05, seed=4401)
    for name, b, want_launches in (("clean", clean, 2), ("with N", dirty, None)):
        want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_START)
        with Aligner(algo=capi.LOCAL, start=capi.WITH_START, max_query_len=160, max_target_len=312, max_n_alns=b.n, n_streams=1) as al:
            before = al.launches(0)
            got = al.align(b)
            n_launches = al.launches(0) - before
            assert al.kernel_name(0) == "s16x2"
            assert_same(got, want, K5, name)
            if want_launches is not None:
                assert n_launches == want_launches, n_launches                                   # counter reset + the DP kernel
            else:
                assert n_launches >= 5
            # one sequence at a time (gasal_host_batch_fill pads by itself): same verdict, same results
            al.vec.a[0].current_n_alns = 0
            al.fill_one_by_one(0, [b.query(i) for i in range(b.n)], [b.target(i) for i in range(b.n)])
            qb, tb = al._last_bytes
            al.launch(0, qb, tb, b.n); al.wait(0)
            assert_same(al.results(0, b.n), want, K5, name + ", fill one by one")
    lib.gasalx_set_fast_path_passes(1)
    try:
        with Aligner(algo=capi.SEMI_GLOBAL, start=capi.WITH_START, max_query_len=160, max_target_len=312, max_n_alns=clean.n, n_streams=1) as al:
            before = al.launches(0)
            got = al.align(clean)
            assert al.launches(0) - before >= 5
        assert_same(got, oracle_align(oracle, clean, algo=po.SEMI_GLOBAL, start=po.WITH_START, max_query_len=1024), K5, "forced passes")
    finally:
        lib.gasalx_set_fast_path_passes(0)
