"""GPU parity of the full-matrix NW kernel (csrc/nw_matrix.cu, through the C ABI of include/nw_b200.h) against
oracle/nw_linear.c -- bit-exact on every cell -- plus the full-size N = 32000 run checked through the matrix checksum."""
import os

import numpy as np
import pytest

from genomics_gpu_b200 import nw as nwmod, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nw_golden.npz")


@pytest.fixture(scope="module")
def orc():
    return po.NwOracle()


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 127, 128, 129, 255, 256, 260, 1000, 1023, 4100])
def test_matrix_bit_exact(orc, n, variant):
    """variant 0: warp-specialised kernel (default); 1: one warp per row-block"""
    ref, qry = synth.nw_benchmark_pair(n, seed=100 + n)
    nwmod.set_variant(variant)
    try:
        got, ms, h = nwmod.nw_matrix(ref, qry, want_checksum=True)
    finally:
        nwmod.set_variant(0)
    want = orc.matrix(ref, qry)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"N={n}: {bad.shape[0]} cells differ, first {bad[:3].tolist()}"
    assert h == po.nw_checksum_of(want)


def test_golden_vectors_from_the_reference_build():
    g = np.load(GOLD)
    for n in g["sizes"]:
        got, _, h = nwmod.nw_matrix(g[f"ref_{n}"], g[f"qry_{n}"], want_checksum=True)
        assert np.array_equal(got, g[f"mat_{n}"]), n
        assert h == int(g[f"sum_{n}"])


@pytest.mark.parametrize("scores", [(2, -3, -2, -1), (5, -4, -7, -7), (3, 0, -1, -5)])
def test_general_scores(orc, scores):
    ref, qry = synth.nw_benchmark_pair(777, seed=5)
    got, _, _ = nwmod.nw_matrix(ref, qry, scores)
    assert np.array_equal(got, orc.matrix(ref, qry, scores))


def test_degenerate_inputs(orc):
    z = np.zeros(0, dtype=np.uint8)
    got, ms, h = nwmod.nw_matrix(z, z, want_checksum=True)
    assert got.size == 0 and h == 0
    same = np.full(500, 65, dtype=np.uint8)                    # all symbols equal: every diagonal extends
    got, _, _ = nwmod.nw_matrix(same, same)
    assert np.array_equal(got, orc.matrix(same, same)) and got[-1, -1] == 500
    a = np.full(300, 1, dtype=np.uint8); b = np.full(300, 2, dtype=np.uint8)   # nothing matches
    got, _, _ = nwmod.nw_matrix(a, b)
    assert np.array_equal(got, orc.matrix(a, b))


def test_repeated_launches_reuse_the_boundary_buffer(orc):
    """the (value, tag) boundary entries of an earlier launch must never validate in a later one"""
    for seed in range(4):
        ref, qry = synth.nw_benchmark_pair(2048, seed=seed)
        got, _, _ = nwmod.nw_matrix(ref, qry)
        assert np.array_equal(got, orc.matrix(ref, qry)), seed


def test_full_size_32000_checksum(orc):
    """the benchmark's default N (NW/main.cu:67): 1.02e9 cells, 4.1 GB matrix kept on the device, compared with the CPU
    restatement through the streaming checksum; plus the last row copied back through a second, smaller run"""
    n = 32000
    ref, qry = synth.nw_benchmark_pair(n, seed=7)
    _, ms, h = nwmod.nw_matrix(ref, qry, want_matrix=False, want_checksum=True)
    assert h == orc.checksum(ref, qry)
    nwmod.release()
