"""CPU tests of the host-side 4-bit packer (genomics_gpu_b200/csrc/host_pack.cpp) that gasal_host_batch_fill / _add use when host
packing is on: same layout as the reference's device kernel (Non-CDP/GASAL2/src/kernels/pack_rc_seqs.h:13-53), checked against
the oracle's restatement of that kernel and the survey's known answer."""
import ctypes as C

import numpy as np

from genomics_gpu_b200 import capi
from oracle import pyoracle as po


def ref_pack(ascii_bytes: np.ndarray) -> np.ndarray:
    a = (ascii_bytes.astype(np.uint32) & 15).reshape(-1, 8)
    sh = np.array([28, 24, 20, 16, 12, 8, 4, 0], dtype=np.uint32)
    return (a << sh).sum(axis=1).astype(np.uint32)


def test_known_answer_and_oracle():
    lib = capi.load()
    a = np.frombuffer(b"ACGTNacg", dtype=np.uint8).copy()
    out = np.zeros(1, dtype=np.uint32)
    lib.gasalx_pack_host(a.ctypes.data, 8, out.ctypes.data)
    assert out[0] == 0x1374E137                                       # SURVEY probe (pack_rc_seqs.h:13-53)
    rng = np.random.default_rng(5)
    for n in (8, 16, 24, 32, 40, 4096, 8 * 100_003):                  # below / above the 32-base SIMD step, odd word counts
        a = rng.integers(0, 256, n, dtype=np.uint8)
        out = np.zeros(n // 8, dtype=np.uint32)
        lib.gasalx_pack_host(a.ctypes.data, n, out.ctypes.data)
        assert np.array_equal(out, ref_pack(a))
        if n <= 4096:
            assert np.array_equal(out, po.Oracle().pack(a))


def test_pack_at_any_alignment_keeps_neighbours():
    lib = capi.load()
    rng = np.random.default_rng(6)
    total = 4096
    for _ in range(200):
        base = rng.integers(0, 256, total, dtype=np.uint8)
        words = ref_pack(base).copy()
        pos = int(rng.integers(0, total - 1)); n = int(rng.integers(0, min(600, total - pos)))
        new = rng.integers(0, 256, max(n, 1), dtype=np.uint8)[:n]
        lib.gasalx_pack_host_at(words.ctypes.data, pos, new.ctypes.data if n else None, n)
        want = base.copy(); want[pos:pos + n] = new
        assert np.array_equal(words, ref_pack(want)), (pos, n)
