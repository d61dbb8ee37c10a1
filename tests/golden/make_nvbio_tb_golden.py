"""Generates tests/golden/nvbio_tb_golden.npz from the REFERENCE's own host code path (oracle/_ref/libnvbio_ref.so =
Non-CDP/NvB/nvbio/alignment headers compiled in place by `make -C oracle ref_nvbio`): banded scores
(aln::banded_alignment_score<BAND_LEN>), full-matrix tracebacks (aln::alignment_traceback) and banded tracebacks
(aln::banded_alignment_traceback<BAND_LEN>).  Needs /root/reference; run once:
    python tests/golden/make_nvbio_tb_golden.py
Per (aligner, scheme, type): rows of [score, source text, source pattern, sink text, sink pattern, n_ops, clip_end, clip_begin]
and the concatenated edit operations in push order; banded scores as [score, sink text, sink pattern, ok]."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from genomics_gpu_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

CONFIGS = [(po.NVB_GOTOH, (2, -1, -2, -1)), (po.NVB_GOTOH, (0, -5, -8, -3)), (po.NVB_SW, (2, -1, -1, -2)), (po.NVB_ED, (0, 0, 0, 0))]
TYPES = (po.NVB_GLOBAL, po.NVB_LOCAL, po.NVB_SEMI_GLOBAL)
SCORE_BANDS = (3, 5, 7, 15, 31)
TB_BANDS = (7, 15, 31)


def tb_rows(orc, algo, type_, scheme, jobs, band=0):
    rows, ops = [], []
    for p, t in jobs:
        r = orc.traceback(algo, type_, scheme, p, t, band=band, ref=True)
        rows.append([r["score"], *r["source"], *r["sink"], len(r["ops"]), r["clip_end"], r["clip_begin"]])
        ops.append(r["ops"])
    return np.array(rows, np.int32), (np.concatenate(ops) if ops else np.zeros(0, np.uint8))


def main():
    orc = po.NvbioOracle()
    assert orc.ref is not None and hasattr(orc.ref, "nvbio_ref_traceback"), "build oracle/_ref/libnvbio_ref.so first (make -C oracle ref_nvbio)"
    out = {"configs": np.array([[a, *s] for a, s in CONFIGS], np.int32)}
    fp, fpo, ft, fto = synth.nvb_jobs(96, 1, 90, 1, 160, seed=21)
    out.update(full_patterns=fp, full_poffs=fpo, full_texts=ft, full_toffs=fto)
    full_jobs = [(fp[fpo[k]:fpo[k + 1]], ft[fto[k]:fto[k + 1]]) for k in range(len(fpo) - 1)]
    bp, bpo, bt, bto = synth.nvb_banded_jobs(96, 1, 110, 31, seed=22)
    out.update(band_patterns=bp, band_poffs=bpo, band_texts=bt, band_toffs=bto)
    band_jobs = [(bp[bpo[k]:bpo[k + 1]], bt[bto[k]:bto[k + 1]]) for k in range(len(bpo) - 1)]
    for ci, (algo, scheme) in enumerate(CONFIGS):
        for type_ in TYPES:
            rows, ops = tb_rows(orc, algo, type_, scheme, full_jobs)
            out[f"full_{ci}_{type_}"] = rows; out[f"full_ops_{ci}_{type_}"] = ops
            for band in SCORE_BANDS:
                out[f"bscore_{ci}_{type_}_{band}"] = np.array([orc.banded_score(algo, type_, band, scheme, p, t, ref=True) for p, t in band_jobs], np.int32)
            for band in TB_BANDS:
                rows, ops = tb_rows(orc, algo, type_, scheme, band_jobs, band=band)
                out[f"btb_{ci}_{type_}_{band}"] = rows; out[f"btb_ops_{ci}_{type_}_{band}"] = ops
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nvbio_tb_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
