"""Generates tests/golden/nvbio_golden.npz from the REFERENCE's own host code path (oracle/_ref/libnvbio_ref.so =
Non-CDP/NvB/nvbio/alignment headers compiled in place by `make -C oracle ref_nvbio`).  Needs /root/reference; run once:
    python tests/golden/make_nvbio_golden.py
The fixture holds the jobs and, per (aligner, type, scheme), BestSink score / text end / pattern end of every job."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from genomics_gpu_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

CONFIGS = [  # (algo, scheme): sw-benchmark's Gotoh scheme (sw-benchmark.cu:585-589), a BWA-like one, SW, edit distance
    (po.NVB_GOTOH, (2, -1, -2, -1)), (po.NVB_GOTOH, (1, -4, -7, -1)), (po.NVB_SW, (2, -1, -2, -2)), (po.NVB_SW, (1, -2, -3, -1)), (po.NVB_ED, (0, 0, 0, 0))]


def main():
    orc = po.NvbioOracle()
    assert orc.ref is not None, "build oracle/_ref/libnvbio_ref.so first (make -C oracle ref_nvbio)"
    pats, poffs, texts, toffs = synth.nvb_jobs(240, 1, 90, 1, 200, seed=11)
    out = {"patterns": pats, "poffs": poffs, "texts": texts, "toffs": toffs, "configs": np.array([[a, *s] for a, s in CONFIGS], np.int32)}
    for ci, (algo, scheme) in enumerate(CONFIGS):
        for type_ in (po.NVB_GLOBAL, po.NVB_LOCAL, po.NVB_SEMI_GLOBAL):
            res = np.array([orc.ref_score(algo, type_, scheme, pats[poffs[k]:poffs[k + 1]], texts[toffs[k]:toffs[k + 1]]) for k in range(len(poffs) - 1)], np.int32)
            out[f"res_{ci}_{type_}"] = res
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nvbio_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
