"""Seeded synthetic read/reference-window pairs for the BASELINE.json configs.

The layout is exactly what a GASAL2 caller builds with gasal_host_batch_fill
(reference Non-CDP/GASAL2/src/host_batch.cpp:79-153, test_prog/test_prog.cpp:279-315):
concatenated ASCII, every sequence padded with 'N' to a multiple of 8 bytes, `offsets` = byte offset
of each (padded) sequence, `lens` = true lengths.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence, Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
N_CODE = 0x4E  # 'N'


@dataclass
class PairBatch:
    q_bytes: np.ndarray   # uint8, padded ASCII, concatenated
    t_bytes: np.ndarray
    q_offs: np.ndarray    # uint32 byte offsets (multiples of 8)
    t_offs: np.ndarray
    q_lens: np.ndarray    # uint32 true lengths
    t_lens: np.ndarray

    @property
    def n(self) -> int:
        return int(self.q_lens.size)

    @property
    def cells(self) -> int:
        """GCUPS numerator: sum of true query_len * target_len."""
        return int((self.q_lens.astype(np.int64) * self.t_lens.astype(np.int64)).sum())

    def slice(self, lo: int, hi: int) -> "PairBatch":
        """Pairs [lo, hi) re-based to offset 0 (sequences are stored in pair order)."""
        qo, to = int(self.q_offs[lo]), int(self.t_offs[lo])
        qe = int(self.q_offs[hi - 1]) + _pad8(int(self.q_lens[hi - 1]))
        te = int(self.t_offs[hi - 1]) + _pad8(int(self.t_lens[hi - 1]))
        return PairBatch(self.q_bytes[qo:qe], self.t_bytes[to:te], (self.q_offs[lo:hi] - qo).astype(np.uint32),
                         (self.t_offs[lo:hi] - to).astype(np.uint32), self.q_lens[lo:hi].copy(), self.t_lens[lo:hi].copy())

    def query(self, i: int) -> bytes:
        o = int(self.q_offs[i]); return self.q_bytes[o:o + int(self.q_lens[i])].tobytes()

    def target(self, i: int) -> bytes:
        o = int(self.t_offs[i]); return self.t_bytes[o:o + int(self.t_lens[i])].tobytes()


def _pad8(n: int) -> int:
    return (n + 7) & ~7


def batch_from_sequences(queries: Sequence[bytes], targets: Sequence[bytes]) -> PairBatch:
    """gasal_host_batch_fill semantics for a list of (query, target) byte strings."""
    assert len(queries) == len(targets)

    def cat(seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        padded = ((lens + 7) // 8 * 8).astype(np.int64)
        offs = np.zeros(len(seqs), dtype=np.int64)
        if len(seqs) > 1:
            offs[1:] = np.cumsum(padded)[:-1]
        buf = np.full(int(padded.sum()), N_CODE, dtype=np.uint8)
        for s, o in zip(seqs, offs):
            buf[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
        return buf, offs.astype(np.uint32), lens

    qb, qo, ql = cat(queries)
    tb, to, tl = cat(targets)
    return PairBatch(qb, tb, qo, to, ql, tl)


def _uniform_batch(q: np.ndarray, t: np.ndarray) -> PairBatch:
    """q: [n, Q] uint8 ASCII, t: [n, T] uint8 ASCII, all pairs the same shape."""
    n, Q = q.shape
    _, T = t.shape
    Q8, T8 = _pad8(Q), _pad8(T)
    qb = np.full((n, Q8), N_CODE, dtype=np.uint8); qb[:, :Q] = q
    tb = np.full((n, T8), N_CODE, dtype=np.uint8); tb[:, :T] = t
    return PairBatch(qb.reshape(-1), tb.reshape(-1), (np.arange(n, dtype=np.uint64) * Q8).astype(np.uint32),
                     (np.arange(n, dtype=np.uint64) * T8).astype(np.uint32), np.full(n, Q, dtype=np.uint32), np.full(n, T, dtype=np.uint32))


def _mutated_windows(rng: np.random.Generator, targets: np.ndarray, starts: np.ndarray, qlen: int, sub: float, indel: float,
                     max_indel: int) -> np.ndarray:
    """query[i, j] = targets[i, starts[i] + walk(j)] with substitutions and short indels (vectorised)."""
    n, T = targets.shape
    step = np.ones((n, qlen), dtype=np.int64)
    ins = np.zeros((n, qlen), dtype=bool)
    if indel > 0:
        ev = rng.random((n, qlen))
        is_ins = ev < indel / 2
        is_del = (ev >= indel / 2) & (ev < indel)
        k = rng.integers(1, max_indel + 1, size=(n, qlen))
        step = np.where(is_del, 1 + k, step)
        step = np.where(is_ins, 0, step)
        ins = is_ins
    pos = starts[:, None] + np.cumsum(step, axis=1) - step   # source index of each query base
    pos = np.clip(pos, 0, T - 1)
    q = np.take_along_axis(targets, pos, axis=1)
    rnd = ACGT[rng.integers(0, 4, size=(n, qlen))]
    q = np.where(ins, rnd, q)
    if sub > 0:
        m = rng.random((n, qlen)) < sub
        # substitute with a DIFFERENT base
        alt = ACGT[(np.searchsorted(ACGT, q, sorter=np.argsort(ACGT)) * 0 + rng.integers(1, 4, size=(n, qlen)) + _code(q)) % 4]
        q = np.where(m, alt, q)
    return q.astype(np.uint8)


def _code(a: np.ndarray) -> np.ndarray:
    """A,C,G,T -> 0..3"""
    lut = np.zeros(256, dtype=np.int64)
    lut[ord("A")], lut[ord("C")], lut[ord("G")], lut[ord("T")] = 0, 1, 2, 3
    return lut[a]


def make_read_window_pairs(n: int, qlen: int, tlen: int, *, sub: float, indel: float = 0.0, max_indel: int = 3, seed: int = 1) -> PairBatch:
    """C1/C2/C4-style: target = random ACGT window, query = a mutated sub-window of it."""
    rng = np.random.default_rng(seed)
    targets = ACGT[rng.integers(0, 4, size=(n, tlen))]
    slack = max(tlen - qlen - max_indel * 4, 0)
    starts = rng.integers(0, slack + 1, size=n)
    q = _mutated_windows(rng, targets, starts, qlen, sub, indel, max_indel)
    return _uniform_batch(q, targets)


def make_extension_pairs(n: int, qlen: int, tlen: int, *, sub: float = 0.02, indel: float = 0.002, seed: int = 9) -> PairBatch:
    """seed-extension shape (KSW, ksw_kernel_template.h): the query is a mutated copy of the target's FIRST qlen bases, so an extension that
    starts at (0, 0) with a seed score runs the whole query"""
    rng = np.random.default_rng(seed)
    targets = ACGT[rng.integers(0, 4, size=(n, tlen))]
    q = _mutated_windows(rng, targets, np.zeros(n, dtype=np.int64), qlen, sub, indel, 3)
    return _uniform_batch(q, targets)


def config1(n: int = 10_000, seed: int = 1) -> PairBatch:
    """C1: 150 bp query / 300 bp target, 3 % substitutions (local)."""
    return make_read_window_pairs(n, 150, 300, sub=0.03, seed=seed)


def config2(n: int = 1_000_000, seed: int = 2) -> PairBatch:
    """C2: 150 x 600 bp, 2 % subs + 0.3 % 1-3 bp indels (local / semi-global with start)."""
    return make_read_window_pairs(n, 150, 600, sub=0.02, indel=0.003, max_indel=3, seed=seed)


def config4(n: int = 2_000_000, seed: int = 4) -> PairBatch:
    """C4: 250 bp Illumina-like read vs its locus +-24 bp (t = 298), 1 % subs, 0.05 % indels <= 8 bp (banded + CIGAR)."""
    rng = np.random.default_rng(seed)
    tlen, qlen = 298, 250
    targets = ACGT[rng.integers(0, 4, size=(n, tlen))]
    starts = np.full(n, 24, dtype=np.int64)
    q = _mutated_windows(rng, targets, starts, qlen, 0.01, 0.0005, 8)
    return _uniform_batch(q, targets)


def nw_style_pair(rng: np.random.Generator, n: int) -> Tuple[bytes, bytes]:
    """C3: the NW benchmark's input model (reference Non-CDP/NW/main.cu:40-61): the query copies the target
    with probability 0.8, inserts a random base with 0.1, skips a target base with 0.1."""
    t = ACGT[rng.integers(0, 4, size=n)]
    ev = rng.random(2 * n)
    out = np.empty(n, dtype=np.uint8)
    rnd = ACGT[rng.integers(0, 4, size=2 * n)]
    ti = 0; qi = 0; k = 0
    # vectorising this walk is not worth it: n <= 16384 and only a few hundred pairs are ever made
    while qi < n:
        e = ev[k]
        if e < 0.8 and ti < n:
            out[qi] = t[ti]; ti += 1; qi += 1
        elif e < 0.9 or ti >= n:
            out[qi] = rnd[k]; qi += 1
        else:
            ti += 1
        k = (k + 1) % (2 * n)
    return out.tobytes(), t.tobytes()


def nw_benchmark_pair(n: int, seed: int = 7) -> Tuple[np.ndarray, np.ndarray]:
    """The NW benchmark's own input model (Non-CDP/NW/main.cu:40-61,93-96): reference bytes uniform in 0..255, the query
    copies the reference with probability 0.8, inserts a random byte with 0.1, skips a reference byte with 0.1.
    Vectorised: the walk is a cumulative sum of the per-event query / reference advances.  Returns (reference, query)."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 256, size=n, dtype=np.uint8)
    m = 2 * n + 64
    ev = rng.random(m)
    copy, ins = ev < 0.8, (ev >= 0.8) & (ev < 0.9)
    r_adv = np.cumsum(copy | ~(copy | ins)); q_adv = np.cumsum(copy | ins)     # reference / query positions consumed AFTER each event
    r_before = r_adv - (copy | ~(copy | ins)); q_before = q_adv - (copy | ins)
    ok = (r_before < n) & (q_before < n)                                       # main.cu:48 `while (r < N && q < N)`
    rnd = rng.integers(0, 256, size=m, dtype=np.uint8)
    qry = rng.integers(0, 256, size=n, dtype=np.uint8)                          # tail insertions (main.cu:58-60)
    sel = ok & copy
    qry[q_before[sel]] = ref[r_before[sel]]
    sel = ok & ins
    qry[q_before[sel]] = rnd[sel]
    return ref, qry


def config3(n_pairs: int = 128, length: int = 16_384, seed: int = 3) -> PairBatch:
    rng = np.random.default_rng(seed)
    qs, ts = zip(*[nw_style_pair(rng, length) for _ in range(n_pairs)])
    return batch_from_sequences(list(qs), list(ts))


def config5(n: int = 200_000, seed: int = 5, lo: int = 100, hi: int = 10_000) -> PairBatch:
    """C5: long-read-like mixed lengths: q log-uniform in [lo, hi], t = q * U[1.0, 1.3], 8 % subs + 4 % indels."""
    rng = np.random.default_rng(seed)
    qlens = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)
    tlens = (qlens * rng.uniform(1.0, 1.3, size=n)).astype(np.int64)
    queries: List[bytes] = []
    targets: List[bytes] = []
    for ql, tl in zip(qlens, tlens):
        t = ACGT[rng.integers(0, 4, size=(1, int(tl)))]
        start = rng.integers(0, max(int(tl - ql), 0) + 1, size=1)
        q = _mutated_windows(rng, t, start, int(ql), 0.08, 0.04, 3)
        queries.append(q.tobytes()); targets.append(t.tobytes())
    return batch_from_sequences(queries, targets)


def random_ragged_pairs(n: int, qmin: int, qmax: int, tmin: int, tmax: int, *, seed: int, related: float = 0.7, sub: float = 0.05,
                        indel: float = 0.03, n_frac: float = 0.0, repeat_frac: float = 0.2) -> PairBatch:
    """Parity-test generator: ragged lengths, a mix of related and unrelated pairs, optional real N bases,
    and tandem-repeat pairs that force ties in the argmax."""
    rng = np.random.default_rng(seed)
    queries: List[bytes] = []
    targets: List[bytes] = []
    for _ in range(n):
        ql = int(rng.integers(qmin, qmax + 1)); tl = int(rng.integers(tmin, tmax + 1))
        t = ACGT[rng.integers(0, 4, size=(1, tl))]
        if rng.random() < repeat_frac and tl >= 24:
            unit = ACGT[rng.integers(0, 4, size=int(rng.integers(4, 12)))]
            t = np.resize(unit, tl)[None, :].copy()
            if rng.random() < 0.5:
                t[0, rng.integers(0, tl, size=max(tl // 20, 1))] = ACGT[rng.integers(0, 4, size=max(tl // 20, 1))]
        if rng.random() < related and tl > 8:
            start = rng.integers(0, max(tl - ql, 0) + 1, size=1)
            q = _mutated_windows(rng, t, start, ql, sub, indel, 3)
        else:
            q = ACGT[rng.integers(0, 4, size=(1, ql))]
        q = q.reshape(-1).copy(); tt = t.reshape(-1).copy()
        if n_frac > 0:
            q[rng.random(ql) < n_frac] = N_CODE
            tt[rng.random(tl) < n_frac] = N_CODE
        queries.append(q.tobytes()); targets.append(tt.tobytes())
    return batch_from_sequences(queries, targets)


# ---- nvbio-style batched scoring jobs (include/nvbio_b200.h; NvB/sw-benchmark shape: reads against text windows) ----------
def nvb_jobs(n: int, m_lo: int, m_hi: int, n_lo: int, n_hi: int, seed: int = 1, sub: float = 0.05, indel: float = 0.02, shared_text: bool = False):
    """n (pattern, text) jobs over codes 0..3: text random, pattern = a window of its text with substitutions and short indels
    (a third of the patterns are unrelated random strings).  Returns (patterns, pattern_offsets, texts, text_offsets)."""
    rng = np.random.default_rng(seed)
    n_texts = 1 if shared_text else n
    t_lens = rng.integers(n_lo, n_hi + 1, n_texts)
    texts = [rng.integers(0, 4, int(L)).astype(np.uint8) for L in t_lens]
    pats = []
    for k in range(n):
        t = texts[0 if shared_text else k]
        m = int(rng.integers(m_lo, m_hi + 1))
        if k % 3 == 2:
            p = rng.integers(0, 4, m).astype(np.uint8)
        else:
            a = int(rng.integers(0, max(1, len(t) - m + 1)))
            p = t[a:a + m].copy()
            flip = rng.random(len(p)) < sub
            p[flip] = rng.integers(0, 4, int(flip.sum()))
            if indel > 0 and len(p) > 4:
                for _ in range(int(rng.binomial(len(p), indel))):
                    pos = int(rng.integers(1, len(p) - 1))
                    if rng.random() < 0.5:
                        p = np.delete(p, pos)
                    else:
                        p = np.insert(p, pos, rng.integers(0, 4, int(rng.integers(1, 4)))).astype(np.uint8)
            if len(p) == 0:
                p = rng.integers(0, 4, 1).astype(np.uint8)
        pats.append(np.ascontiguousarray(p, dtype=np.uint8))
    poffs = np.zeros(n + 1, np.uint32); poffs[1:] = np.cumsum([len(p) for p in pats])
    toffs = np.zeros(n_texts + 1, np.uint32); toffs[1:] = np.cumsum([len(t) for t in texts])
    return np.concatenate(pats), poffs, np.concatenate(texts), toffs


def long_read_windows(n: int, lo: int, hi: int, flank: int, seed: int = 5) -> PairBatch:
    """n long reads (length uniform in [lo, hi), NW-benchmark style errors: 80 % copy / 10 % insert / 10 % delete) against their source
    window extended by up to `flank` random bases on either side: the semi-global read-vs-window shape for long reads"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    qs, ts = [], []
    for _ in range(n):
        t, q = nw_style_pair(rng, int(rng.integers(lo, hi)))
        qs.append(q)
        ts.append(bytes(rng.choice(acgt, size=int(rng.integers(0, flank + 1)))) + t + bytes(rng.choice(acgt, size=int(rng.integers(0, flank + 1)))))
    return batch_from_sequences(qs, ts)


def nvb_banded_jobs(n: int, m_lo: int, m_hi: int, band: int, seed: int = 1, sub: float = 0.05, indel: float = 0.02, slack: int = 6):
    """n jobs for the banded aligners (row = pattern symbol, band entry j = text symbol i + j): text = pattern length + band - 1 + up to
    `slack` symbols (the reference's banded-traceback domain), pattern = the text from an offset inside the band, with substitutions and
    short indels; every seventh pattern is unrelated.  Codes 0..3.  Returns (patterns, pattern_offsets, texts, text_offsets)."""
    rng = np.random.default_rng(seed)
    pats, texts = [], []
    for k in range(n):
        m = int(rng.integers(m_lo, m_hi + 1))
        t = rng.integers(0, 4, m + band - 1 + int(rng.integers(0, slack + 1)) + 4).astype(np.uint8)
        if k % 7 == 6:
            p = rng.integers(0, 4, m).astype(np.uint8)
        else:
            a = int(rng.integers(0, band))
            p = t[a:a + m].copy()
            flip = rng.random(len(p)) < sub
            p[flip] = rng.integers(0, 4, int(flip.sum()))
            if indel > 0 and len(p) > 4:
                for _ in range(int(rng.binomial(len(p), indel))):
                    pos = int(rng.integers(1, len(p) - 1))
                    p = np.delete(p, pos) if rng.random() < 0.5 else np.insert(p, pos, rng.integers(0, 4, int(rng.integers(1, 3)))).astype(np.uint8)
            p = p[:m] if len(p) > m else p
            if len(p) == 0:
                p = rng.integers(0, 4, 1).astype(np.uint8)
        t = t[:max(len(t) - 4, len(p) + band - 1)]
        pats.append(np.ascontiguousarray(p, dtype=np.uint8)); texts.append(np.ascontiguousarray(t, dtype=np.uint8))
    poffs = np.zeros(n + 1, np.uint32); poffs[1:] = np.cumsum([len(p) for p in pats])
    toffs = np.zeros(n + 1, np.uint32); toffs[1:] = np.cumsum([len(t) for t in texts])
    return np.concatenate(pats), poffs, np.concatenate(texts), toffs


def with_n_bases(batch: PairBatch, pair_frac: float, seed: int = 1, per_pair: int = 1) -> PairBatch:
    """Copy of `batch` in which a fraction of the pairs carries `per_pair` real 'N' bases (inside the true length, in the query or
    the target at random): reads with uncalled bases, the input the s16x2 fast path has to hand to its N-aware second pass."""
    rng = np.random.default_rng(seed)
    q = batch.q_bytes.copy(); t = batch.t_bytes.copy()
    hit = np.nonzero(rng.random(batch.n) < pair_frac)[0]
    for _ in range(per_pair):
        in_q = rng.random(hit.size) < 0.5
        hq, ht = hit[in_q], hit[~in_q]
        q[batch.q_offs[hq].astype(np.int64) + (rng.random(hq.size) * batch.q_lens[hq]).astype(np.int64)] = N_CODE
        t[batch.t_offs[ht].astype(np.int64) + (rng.random(ht.size) * batch.t_lens[ht]).astype(np.int64)] = N_CODE
    return PairBatch(q, t, batch.q_offs, batch.t_offs, batch.q_lens, batch.t_lens)


def overhanging_reads(n: int, seed: int, qlen: int = 150, tlen: int = 400) -> PairBatch:
    """reads that hang off either end of their window, start on its first bases, or carry long gaps: the shapes where the pruned start
    pass (score bound on the reversed rectangle) must fall back or must keep the far corner"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    qs, ts = [], []
    for i in range(n):
        g = acgt[rng.integers(0, 4, tlen + 2 * qlen)]
        t = g[qlen:qlen + tlen]
        kind = i % 6
        if kind == 0: s = qlen - int(rng.integers(1, 40))                      # first bases of the read lie before the window
        elif kind == 1: s = qlen + tlen - qlen + int(rng.integers(1, 40))       # last bases lie behind it
        elif kind == 2: s = qlen + int(rng.integers(0, 3))                      # starts on the first rows
        elif kind == 3: s = qlen + tlen - qlen - int(rng.integers(0, 3))        # ends on the last rows
        else: s = qlen + int(rng.integers(0, tlen - qlen))
        q = g[s:s + qlen].copy()
        if kind == 4:                                                          # a long deletion / insertion inside the read
            cut = int(rng.integers(20, 120)); L = int(rng.integers(5, 30))
            q = np.concatenate([q[:cut], q[cut + L:], acgt[rng.integers(0, 4, L)]])[:qlen]
        if kind == 5:
            cut = int(rng.integers(20, 120)); L = int(rng.integers(5, 30))
            q = np.concatenate([q[:cut], acgt[rng.integers(0, 4, L)], q[cut:]])[:qlen]
        sub = rng.random(q.size) < 0.02
        q[sub] = acgt[rng.integers(0, 4, int(sub.sum()))]
        qs.append(q.tobytes()); ts.append(t.tobytes())
    return batch_from_sequences(qs, ts)
