"""Shared helpers for the parity tests: run one batch through the oracle and compare field by field."""
from __future__ import annotations

import numpy as np

from oracle import pyoracle as po

K3 = ["aln_score", "query_batch_end", "target_batch_end"]
K5 = K3 + ["query_batch_start", "target_batch_start"]
K2ND = ["aln_score2", "query_batch_end2", "target_batch_end2"]


def oracle_align(orc, batch, **kw):
    pq = orc.pack(batch.q_bytes)
    pt = orc.pack(batch.t_bytes)
    return orc.align(pq, pt, batch.q_lens, batch.t_lens, batch.q_offs, batch.t_offs, **kw)


def assert_same(got, want, keys, what=""):
    for k in keys:
        assert k in got, f"{what}: engine did not return {k}"
        g, w = np.asarray(got[k]), np.asarray(want[k])
        bad = np.nonzero(g != w)[0]
        assert bad.size == 0, f"{what}: {k} differs for {bad.size}/{g.size} pairs, first idx {bad[:5]}: got {g[bad[:5]]}, want {w[bad[:5]]}"


def cigar_slices(res, batch):
    """Per-pair CIGAR byte strings (reverse alignment order, as stored)."""
    out = []
    for i in range(batch.n):
        o = int(batch.q_offs[i]); n = int(res["n_cigar_ops"][i])
        out.append(bytes(res["cigar"][o:o + n]))
    return out


def cigar_rescore(cig: bytes, scoring):
    """Re-score a stored CIGAR (reverse order does not matter for the score)."""
    m, x, o, e = scoring
    s = 0
    prev = -1
    for b in cig:
        op, cnt = b & 3, b >> 2
        if op == 0: s += m * cnt
        elif op == 1: s -= x * cnt
        else:
            s -= e * cnt + (o if prev != op else 0)
        prev = op
    return s
