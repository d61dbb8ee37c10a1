"""GPU parity of the nvbio-style batched scoring adapter (include/nvbio_b200.h) against oracle/nvbio_compat.c -- itself pinned to
the reference's host build of Non-CDP/NvB/nvbio/alignment (tests/test_nvbio_oracle.py) -- and against the committed golden
vectors of that build.  Every call goes through the C ABI."""
import os

import numpy as np
import pytest

from genomics_gpu_b200 import nvb, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nvbio_golden.npz")
TYPES = (nvb.GLOBAL, nvb.LOCAL, nvb.SEMI_GLOBAL)
GOTOH_SW = (2, -1, -2, -1)          # sw-benchmark.cu:585-589


@pytest.fixture(scope="module")
def orc():
    return po.NvbioOracle()


FAST = ("s16x2", "s16x2-strip")     # the band-blocked packed kernel (k_nvb_s16) / the quad-strip kernel (k_nvb_strip, nvb_strip.cuh)
_POLICY = [0]


@pytest.fixture(autouse=True, params=[0, 2], ids=["auto", "no-strip"])
def _fast_policy(request):
    """every test of this module runs twice: with the automatic kernel choice (quad-strip kernel where it is eligible) and with the
    quad-strip kernel switched off (policy 2), so that both packed kernels keep their parity coverage"""
    _POLICY[0] = request.param
    yield
    _POLICY[0] = 0
    nvb.set_policy(0)


def run(al, batch, want_sinks=False, policy=None):
    policy = _POLICY[0] if policy is None else policy
    nvb.set_policy(policy)
    try:
        with nvb.Stream(batch) as st:
            st.enact(al, want_sinks)
            s, si, sj = st.results(want_sinks)
            return s, si, sj, st.kernel()
    finally:
        nvb.set_policy(0)


def test_golden_vectors_through_the_c_abi():
    g = np.load(GOLD)
    batch = nvb.Batch(g["patterns"], g["poffs"], g["texts"], g["toffs"])
    for ci, cfg in enumerate(g["configs"]):
        al_args = [int(x) for x in cfg]
        for type_ in TYPES:
            want = g[f"res_{ci}_{type_}"]
            al = nvb.aligner(al_args[0], type_, *al_args[1:])
            s, si, sj, _ = run(al, batch, want_sinks=True)
            assert np.array_equal(s, want[:, 0]) and np.array_equal(si, want[:, 1]) and np.array_equal(sj, want[:, 2]), (ci, type_)
            s2, _, _, _ = run(al, batch)                       # score-only fast path where eligible
            assert np.array_equal(s2, want[:, 0]), (ci, type_, "fast")


@pytest.mark.parametrize("type_", TYPES)
@pytest.mark.parametrize("scheme", [GOTOH_SW, (1, -4, -7, -1), (3, -2, -5, -5)])
def test_fast_path_scores_ragged_lengths(orc, type_, scheme):
    pats, poffs, texts, toffs = synth.nvb_jobs(3001, 1, 210, 1, 330, seed=21)       # odd count: the last thread holds one job
    batch = nvb.Batch(pats, poffs, texts, toffs)
    want, wi, wj = orc.batch(po.NVB_GOTOH, type_, scheme, pats, poffs, texts, toffs)
    s, _, _, k = run(nvb.aligner(nvb.GOTOH, type_, *scheme), batch)
    assert k in FAST
    assert np.array_equal(s, want)
    s, si, sj, k = run(nvb.aligner(nvb.GOTOH, type_, *scheme), batch, want_sinks=True)
    assert k in FAST                                               # LOCAL sinks: packed keys (scores here stay below 2048)
    assert np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj)
    s, si, sj, k = run(nvb.aligner(nvb.GOTOH, type_, *scheme), batch, want_sinks=True, policy=1)
    assert k == "int32" and np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj)


@pytest.mark.parametrize("type_", TYPES)
def test_smith_waterman_and_edit_distance(orc, type_):
    pats, poffs, texts, toffs = synth.nvb_jobs(1500, 5, 120, 5, 260, seed=22)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    for algo, scheme, fast in ((nvb.SMITH_WATERMAN, (2, -1, -2, -2), True), (nvb.SMITH_WATERMAN, (1, -2, -3, -1), False), (nvb.EDIT_DISTANCE, (0, 0, 0, 0), True)):
        want, wi, wj = orc.batch(algo, type_, scheme, pats, poffs, texts, toffs)
        s, _, _, k = run(nvb.aligner(algo, type_, *scheme), batch)
        assert (k in FAST) if fast else (k == "int32"), (algo, scheme, k)      # asymmetric insertion / deletion costs: general kernel
        assert np.array_equal(s, want), (algo, scheme)
        s, si, sj, _ = run(nvb.aligner(algo, type_, *scheme), batch, want_sinks=True)
        assert np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj), (algo, scheme)


@pytest.mark.parametrize("type_", TYPES)
def test_symbols_outside_acgt_go_to_the_general_kernel(orc, type_):
    pats, poffs, texts, toffs = synth.nvb_jobs(2000, 10, 100, 50, 200, seed=23)
    rng = np.random.default_rng(5)
    pats = pats.copy(); texts = texts.copy()
    pats[rng.integers(0, pats.size, 300)] = 4                  # N of nvbio's DNA_N alphabet
    texts[rng.integers(0, texts.size, 300)] = rng.integers(4, 9, 300)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    want, _, _ = orc.batch(po.NVB_GOTOH, type_, GOTOH_SW, pats, poffs, texts, toffs)
    s, _, _, k = run(nvb.aligner(nvb.GOTOH, type_, *GOTOH_SW), batch)
    assert k in FAST and np.array_equal(s, want)
    # ASCII strings: upper-case ACGT is the fast alphabet, anything else (N, lower case) is compared as raw bytes
    lut = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
    pa, ta = lut[pats], lut[texts]
    want_a, _, _ = orc.batch(po.NVB_GOTOH, type_, GOTOH_SW, pa, poffs, ta, toffs)
    s, _, _, _ = run(nvb.aligner(nvb.GOTOH, type_, *GOTOH_SW), nvb.Batch(pa, poffs, ta, toffs, alphabet=nvb.ASCII))
    assert np.array_equal(s, want_a)


def test_all_against_one_text_and_indexed_pairs(orc):
    pats, poffs, texts, toffs = synth.nvb_jobs(777, 30, 150, 2500, 2500, seed=24, shared_text=True)     # sw-benchmark's shape
    al = nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, *GOTOH_SW)
    want, wi, wj = orc.batch(po.NVB_GOTOH, po.NVB_SEMI_GLOBAL, GOTOH_SW, pats, poffs, texts, toffs)
    s, _, _, k = run(al, nvb.Batch(pats, poffs, texts, toffs))
    assert k in FAST and np.array_equal(s, want)
    s, si, sj, _ = run(al, nvb.Batch(pats, poffs, texts, toffs), want_sinks=True)
    assert np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj)
    rng = np.random.default_rng(9)
    pp = rng.integers(0, 777, 500).astype(np.uint32); pt = np.zeros(500, np.uint32)
    s2, _, _, _ = run(al, nvb.Batch(pats, poffs, texts, toffs, pair_pattern=pp, pair_text=pt))
    assert np.array_equal(s2, want[pp])


def test_range_proof_falls_back_to_int32_for_long_global_jobs(orc):
    pats, poffs, texts, toffs = synth.nvb_jobs(64, 5000, 6000, 6000, 7000, seed=25, sub=0.1, indel=0.01)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    al = nvb.aligner(nvb.GOTOH, nvb.GLOBAL, 1, -4, -7, -3)          # 3 * 13000 > 31000: outside 16 bits
    want, _, _ = orc.batch(po.NVB_GOTOH, po.NVB_GLOBAL, (1, -4, -7, -3), pats, poffs, texts, toffs)
    s, _, _, k = run(al, batch)
    assert k == "int32" and np.array_equal(s, want)
    al = nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1)           # local scores stay below 2 * 6000: fast path
    want, _, _ = orc.batch(po.NVB_GOTOH, po.NVB_LOCAL, (2, -1, -2, -1), pats, poffs, texts, toffs)
    s, _, _, k = run(al, batch)
    assert k in FAST and np.array_equal(s, want)


def test_sw_benchmark_shape_20k_reads(orc):
    pats, poffs, texts, toffs = synth.nvb_jobs(20000, 150, 150, 600, 600, seed=26, sub=0.02, indel=0.003)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    with nvb.Stream(batch) as st:
        assert st.cells() == int(np.sum(np.diff(poffs).astype(np.int64) * np.diff(toffs).astype(np.int64)))
        for type_ in TYPES:
            st.enact(nvb.aligner(nvb.GOTOH, type_, *GOTOH_SW))
            s, _, _ = st.results()
            want, _, _ = orc.batch(po.NVB_GOTOH, type_, GOTOH_SW, pats, poffs, texts, toffs)
            assert np.array_equal(s, want), type_
            assert st.last_ms() > 0
        st.enact(nvb.aligner(nvb.EDIT_DISTANCE, nvb.SEMI_GLOBAL))
        s, _, _ = st.results()
        want, _, _ = orc.batch(po.NVB_ED, po.NVB_SEMI_GLOBAL, (), pats, poffs, texts, toffs)
        assert np.array_equal(s, want)


def test_one_call_host_form_and_errors(orc):
    pats, poffs, texts, toffs = synth.nvb_jobs(300, 10, 80, 40, 160, seed=27)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    s, si, sj, ms = nvb.batched_score(nvb.aligner(nvb.GOTOH, nvb.LOCAL, *GOTOH_SW), batch, want_sinks=True)
    want, wi, wj = orc.batch(po.NVB_GOTOH, po.NVB_LOCAL, GOTOH_SW, pats, poffs, texts, toffs)
    assert np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj) and min(ms) >= 0
    bad = poffs.copy(); bad[5] = bad[4]                               # an empty pattern
    with pytest.raises(RuntimeError):
        nvb.Stream(nvb.Batch(pats, bad, texts, toffs))
    with nvb.Stream(batch) as st:
        with pytest.raises(RuntimeError):
            st.results()                                              # nothing enacted yet
        st.enact(nvb.aligner(nvb.GOTOH, nvb.LOCAL, *GOTOH_SW))
        with pytest.raises(RuntimeError):
            st.results(want_sinks=True)                               # sinks were not computed
        with pytest.raises(RuntimeError):
            st.enact(nvb.aligner(7, nvb.LOCAL))


def test_batches_larger_than_one_launch(orc):
    """270001 jobs against one shared text: the fast path cuts launches of 2 x 131072 jobs (compile-time scratch stride), the int32 path
    cuts by scratch size; job ranges, worklists and the odd last job must line up across launches."""
    n, m = 270001, 12
    rng = np.random.default_rng(33)
    text = rng.integers(0, 4, 1000).astype(np.uint8)
    starts = rng.integers(0, 1000 - m, n)
    pats = text[(starts[:, None] + np.arange(m)[None, :])].astype(np.uint8)
    pats[rng.random(pats.shape) < 0.08] = 3
    pats[::5000, 3] = 4                                                   # a few jobs with N: worklist entries in both launches
    pats = np.ascontiguousarray(pats.reshape(-1)); poffs = (np.arange(n + 1) * m).astype(np.uint32)
    toffs = np.array([0, 1000], np.uint32)
    batch = nvb.Batch(pats, poffs, text, toffs)
    for type_ in (nvb.SEMI_GLOBAL, nvb.LOCAL):
        want, wi, wj = orc.batch(po.NVB_GOTOH, type_, GOTOH_SW, pats, poffs, text, toffs)
        s, _, _, k = run(nvb.aligner(nvb.GOTOH, type_, *GOTOH_SW), batch)
        assert k in FAST and np.array_equal(s, want), type_
    s, si, sj, k = run(nvb.aligner(nvb.GOTOH, nvb.LOCAL, *GOTOH_SW), batch, want_sinks=True)
    assert k in FAST and np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj)
    s, si, sj, k = run(nvb.aligner(nvb.GOTOH, nvb.LOCAL, *GOTOH_SW), batch, want_sinks=True, policy=1)
    assert k == "int32" and np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj)


def test_more_flagged_jobs_than_packed_threads(orc):
    """with the fixed-stride scratch a launch covers 2 x 131072 jobs but the band-border scratch holds one
    column per THREAD of the packed kernel (131072).  When more than 131072 jobs of a launch are flagged (every pattern carries an
    N) and the pattern needs the scratch (M > 16), the one-job-per-thread worklist pass must stay inside the allocation."""
    n, m = 150_000, 40
    rng = np.random.default_rng(34)
    text = rng.integers(0, 4, 300).astype(np.uint8)
    starts = rng.integers(0, 300 - m, n)
    pats = text[(starts[:, None] + np.arange(m)[None, :])].astype(np.uint8)
    pats[rng.random(pats.shape) < 0.05] = 2
    pats[np.arange(n), rng.integers(0, m, n)] = 4                        # one N per pattern: every job goes to the worklist
    pats = np.ascontiguousarray(pats.reshape(-1)); poffs = (np.arange(n + 1) * m).astype(np.uint32)
    toffs = np.array([0, 300], np.uint32)
    batch = nvb.Batch(pats, poffs, text, toffs)
    for type_ in (nvb.SEMI_GLOBAL, nvb.GLOBAL, nvb.LOCAL):
        want, _, _ = orc.batch(po.NVB_GOTOH, type_, GOTOH_SW, pats, poffs, text, toffs)
        s, _, _, k = run(nvb.aligner(nvb.GOTOH, type_, *GOTOH_SW), batch)
        assert k in FAST and np.array_equal(s, want), type_


@pytest.mark.parametrize("algo,scheme", [(nvb.GOTOH, GOTOH_SW), (nvb.SMITH_WATERMAN, (2, -1, -2, -2))])
def test_local_sinks_tie_order_and_key_range(orc, algo, scheme):
    """LOCAL BestSink coordinates on the packed path: repeats make many cells tie, so the reference's report order (bands of 8 for Gotoh, 16
    for Smith-Waterman; last maximum wins) decides; scores above 2047 do not fit the packed keys and go to the int32 kernel."""
    rng = np.random.default_rng(41)
    pats, texts = [], []
    for k in range(1200):
        unit = rng.integers(0, 4, int(rng.integers(2, 9))).astype(np.uint8)          # short tandem repeats: ties everywhere
        t = np.tile(unit, int(rng.integers(6, 40)))[: int(rng.integers(20, 260))]
        p = np.tile(unit, 30)[: int(rng.integers(3, 70))]
        if k % 4 == 0:
            p = rng.integers(0, 4, len(p)).astype(np.uint8)
        if k % 7 == 0:
            p = (3 - t[:min(len(t), 40)]).astype(np.uint8)                           # mostly mismatches: score 0 or tiny
        if k % 11 == 0:                                                              # nothing matches: score 0, the last reported cell (N, M) wins
            p = np.zeros(int(rng.integers(1, 50)), np.uint8); t = rng.integers(1, 4, int(rng.integers(1, 90))).astype(np.uint8)
        pats.append(p); texts.append(t)
    pb, poffs = nvb.concat(pats); tb, toffs = nvb.concat(texts)
    batch = nvb.Batch(pb, poffs, tb, toffs)
    want, wi, wj = orc.batch(algo, po.NVB_LOCAL, scheme, pb, poffs, tb, toffs)
    zero = want == 0
    assert zero.sum() > 50 and np.array_equal(wi[zero], np.diff(toffs)[zero]) and np.array_equal(wj[zero], np.diff(poffs)[zero])
    s, si, sj, k = run(nvb.aligner(algo, nvb.LOCAL, *scheme), batch, want_sinks=True)
    assert k in FAST
    assert np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj)
    # long identical strings: score 2 * 1500 > 2047
    pats, poffs, texts, toffs = synth.nvb_jobs(40, 1400, 1500, 1800, 2000, seed=42, sub=0.01, indel=0.0)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    want, wi, wj = orc.batch(algo, po.NVB_LOCAL, scheme, pats, poffs, texts, toffs)
    s, si, sj, k = run(nvb.aligner(algo, nvb.LOCAL, *scheme), batch, want_sinks=True)
    assert k == "int32" and np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj)
    s, _, _, k = run(nvb.aligner(algo, nvb.LOCAL, *scheme), batch)
    assert k in FAST and np.array_equal(s, want)


def test_cpp_front_end(orc, tmp_path):
    """The reference-named C++ front end (include/nvbio_b200_batched.hpp) through tools/nvb_cpp_example.cpp -- sw-benchmark's
    batch_score_profile call sequence -- against the oracle: Gotoh global / semi-global / local and edit-distance semi-global, scores and sinks."""
    import subprocess
    from test_nvbio_oracle import _write_jobs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "nvb_cpp_example")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build_cpp_example()
    pats, poffs, texts, toffs = synth.nvb_jobs(501, 5, 120, 30, 260, seed=44)
    _write_jobs(tmp_path / "jobs.bin", pats, poffs, texts, toffs)
    r = subprocess.run([exe, str(tmp_path / "jobs.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.array([[int(x) for x in ln.split()] for ln in r.stdout.strip().split("\n")], np.int64)
    assert got.shape == (4 * 501, 4)
    for block, (algo, type_, scheme) in enumerate([(po.NVB_GOTOH, po.NVB_GLOBAL, GOTOH_SW), (po.NVB_GOTOH, po.NVB_SEMI_GLOBAL, GOTOH_SW),
                                                    (po.NVB_GOTOH, po.NVB_LOCAL, GOTOH_SW), (po.NVB_ED, po.NVB_SEMI_GLOBAL, ())]):
        s, si, sj = orc.batch(algo, type_, scheme, pats, poffs, texts, toffs)
        g = got[block * 501:(block + 1) * 501]
        assert np.array_equal(g[:, 1], s) and np.array_equal(g[:, 2], si) and np.array_equal(g[:, 3], sj), (algo, type_)


@pytest.mark.parametrize("type_", TYPES)
def test_quad_strip_kernel(orc, type_):
    """k_nvb_strip (nvb_strip.cuh): K2's blocking -- one job per thread, four text strips in the 16-bit lanes, pattern-side column state in
    shared memory, no band-border scratch -- for scores of all three types and the sinks of GLOBAL / SEMI_GLOBAL.  Ragged lengths
    from 1 symbol up (pad rows and pad columns, texts that end in every lane of a quad-strip), several schemes incl. edit
    distance and zero-cost extension, jobs with N (worklist), the all-against-one text shape, patterns beyond the shared-memory
    limit (falls back to the band-blocked kernel)."""
    for seed, (mlo, mhi, nlo, nhi), scheme in ((51, (1, 210, 1, 330), GOTOH_SW), (52, (1, 40, 1, 70), (1, -4, -7, -1)), (53, (100, 320, 1, 500), (3, -2, -5, -5)),
                                               (54, (5, 150, 5, 700), (2, -3, -4, 0)), (55, (1, 64, 200, 260), (1, -1, -1, -1))):
        pats, poffs, texts, toffs = synth.nvb_jobs(2501, mlo, mhi, nlo, nhi, seed=seed)
        batch = nvb.Batch(pats, poffs, texts, toffs)
        want, wi, wj = orc.batch(po.NVB_GOTOH, type_, scheme, pats, poffs, texts, toffs)
        s, _, _, k = run(nvb.aligner(nvb.GOTOH, type_, *scheme), batch, policy=0)
        assert k == "s16x2-strip", (seed, k)
        assert np.array_equal(s, want), (seed, scheme, np.nonzero(s != want)[0][:5])
        s, si, sj, k = run(nvb.aligner(nvb.GOTOH, type_, *scheme), batch, want_sinks=True, policy=0)
        assert k == ("s16x2" if type_ == nvb.LOCAL and scheme[3] < 0 else "int32" if type_ == nvb.LOCAL else "s16x2-strip"), (seed, k)
        assert np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj), (seed, scheme)
        s, _, _, k = run(nvb.aligner(nvb.GOTOH, type_, *scheme), batch, policy=2)
        assert k != "s16x2-strip" and np.array_equal(s, want), (seed, k)
    # edit distance + symbols outside ACGT + one shared text
    pats, poffs, texts, toffs = synth.nvb_jobs(1500, 20, 150, 900, 900, seed=56, shared_text=True)
    pats = pats.copy(); pats[np.random.default_rng(6).integers(0, pats.size, 200)] = 4
    batch = nvb.Batch(pats, poffs, texts, toffs)
    for algo, oalgo, scheme in ((nvb.EDIT_DISTANCE, po.NVB_ED, ()), (nvb.SMITH_WATERMAN, po.NVB_SW, (2, -1, -2, -2)), (nvb.GOTOH, po.NVB_GOTOH, GOTOH_SW)):
        want, wi, wj = orc.batch(oalgo, type_, scheme, pats, poffs, texts, toffs)
        s, _, _, k = run(nvb.aligner(algo, type_, *scheme), batch, policy=0)
        assert k == "s16x2-strip" and np.array_equal(s, want), (algo, k)
    # a pattern of 353+ symbols does not fit the shared-memory column state
    pats, poffs, texts, toffs = synth.nvb_jobs(200, 300, 420, 300, 500, seed=57)
    want, _, _ = orc.batch(po.NVB_GOTOH, type_, GOTOH_SW, pats, poffs, texts, toffs)
    s, _, _, k = run(nvb.aligner(nvb.GOTOH, type_, *GOTOH_SW), nvb.Batch(pats, poffs, texts, toffs), policy=0)
    assert k == "s16x2" and np.array_equal(s, want)


@pytest.mark.parametrize("algo,oalgo,scheme", [(nvb.GOTOH, po.NVB_GOTOH, GOTOH_SW), (nvb.SMITH_WATERMAN, po.NVB_SW, (2, -1, -2, -2)), (nvb.EDIT_DISTANCE, po.NVB_ED, ())])
def test_min_score_early_exit(orc, algo, oalgo, scheme):
    """the reference's per-job min_score threshold (batched.h:249, alignment.h:456-530; nvbx_stream_set_min_scores): jobs that give up
    keep what their sink held; per-job and uniform thresholds; switching it off restores the fast path.  The restatement it is
    compared with is pinned to the reference's host build (tests/test_nvbio_oracle.py::test_min_score_early_exit_equals_reference_host_path)."""
    if _POLICY[0] != 0:
        pytest.skip("one kernel serves the threshold: nothing to A/B")
    pats, poffs, texts, toffs = synth.nvb_jobs(3000, 1, 200, 1, 300, seed=61)
    pats = pats.copy(); pats[np.random.default_rng(7).integers(0, pats.size, 100)] = 4          # any alphabet
    batch = nvb.Batch(pats, poffs, texts, toffs)
    rng = np.random.default_rng(8)
    with nvb.Stream(batch) as st:
        for type_ in TYPES:
            al = nvb.aligner(algo, type_, *scheme)
            full, _, _ = orc.batch(oalgo, type_, scheme, pats, poffs, texts, toffs)
            per_job = (full + rng.integers(-30, 30, full.size)).astype(np.int32)
            for per, uni in ((per_job, 0), (None, 25), (None, -40), (None, 1 << 20)):
                want, wi, wj = orc.batch_min(oalgo, type_, scheme, pats, poffs, texts, toffs, min_score=uni, min_scores=per)
                st.set_min_scores(per, uni)
                st.enact(al, True)
                s, si, sj = st.results(True)
                assert st.kernel() == "int32-minscore"
                assert np.array_equal(s, want) and np.array_equal(si, wi) and np.array_equal(sj, wj), (type_, uni, np.nonzero(s != want)[0][:5])
                assert (want != full).sum() > 20 or uni == -40 or (algo == nvb.EDIT_DISTANCE and type_ == nvb.LOCAL)     # (edit-distance LOCAL scores are all 0)
            st.set_min_scores(None)                                                             # off again
            st.enact(al)
            s, _, _ = st.results()
            assert st.kernel() != "int32-minscore" and np.array_equal(s, full)
