"""ctypes binding of the nvbio-style batched alignment scoring adapter (include/nvbio_b200.h; replaces
aln::BatchedAlignmentScore<...>::enact of Non-CDP/NvB/nvbio/alignment/batched.h:313-335 as driven by
NvB/sw-benchmark/sw-benchmark.cu:340-444).  No fallback: the functions live in libgasal_b200.so and need a CUDA device."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import capi

GLOBAL, LOCAL, SEMI_GLOBAL = 0, 1, 2                 # aln::AlignmentType (alignment_base.h:54)
GOTOH, SMITH_WATERMAN, EDIT_DISTANCE = 0, 1, 2
CODES, ASCII = 0, 1
NVB_SYMBOLS = ["nvbx_stream_create", "nvbx_stream_enact", "nvbx_stream_results", "nvbx_stream_scores_device", "nvbx_stream_last_ms",
               "nvbx_stream_cells", "nvbx_stream_kernel", "nvbx_stream_destroy", "nvbx_batched_score_host", "nvbx_set_policy",
               "nvbx_kernel_launches", "nvbx_last_error", "nvbx_stream_enact_banded", "nvbx_stream_traceback", "nvbx_stream_ops_stride",
               "nvbx_stream_alignments", "nvbx_stream_set_min_scores"]
OP_SUBSTITUTION, OP_INSERTION, OP_DELETION = 0, 1, 2      # aln::DirectionVector values a Backtracer sees (alignment_base.h:139-142)
BAND_LENS = (3, 5, 7, 15, 31)                             # the reference's BatchedBandedAlignmentScore instantiations


class AlignerT(C.Structure):
    _fields_ = [("algo", C.c_int32), ("type", C.c_int32), ("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_ext", C.c_int32)]


class BatchT(C.Structure):
    _fields_ = [("patterns", C.c_void_p), ("pattern_offsets", C.c_void_p), ("n_patterns", C.c_uint32),
                ("texts", C.c_void_p), ("text_offsets", C.c_void_p), ("n_texts", C.c_uint32),
                ("pair_pattern", C.c_void_p), ("pair_text", C.c_void_p), ("n_pairs", C.c_uint32), ("alphabet", C.c_int32)]


class AlignmentT(C.Structure):
    """nvbx_alignment_t: aln::Alignment<int32> (alignment_base.h:124-136) + what the Backtracer receives"""
    _fields_ = [("score", C.c_int32), ("source_text", C.c_int32), ("source_pattern", C.c_int32), ("sink_text", C.c_int32), ("sink_pattern", C.c_int32),
                ("n_ops", C.c_uint32), ("clip_end", C.c_uint32), ("clip_begin", C.c_uint32)]


_bound = False


def _lib() -> C.CDLL:
    global _bound
    lib = capi.load()
    if not _bound:
        lib.nvbx_stream_create.argtypes = [C.POINTER(BatchT), C.c_void_p]; lib.nvbx_stream_create.restype = C.c_void_p
        lib.nvbx_stream_enact.argtypes = [C.c_void_p, C.POINTER(AlignerT), C.c_int]; lib.nvbx_stream_enact.restype = C.c_int
        lib.nvbx_stream_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]; lib.nvbx_stream_results.restype = C.c_int
        lib.nvbx_stream_scores_device.argtypes = [C.c_void_p]; lib.nvbx_stream_scores_device.restype = C.c_void_p
        lib.nvbx_stream_last_ms.argtypes = [C.c_void_p]; lib.nvbx_stream_last_ms.restype = C.c_float
        lib.nvbx_stream_cells.argtypes = [C.c_void_p]; lib.nvbx_stream_cells.restype = C.c_uint64
        lib.nvbx_stream_kernel.argtypes = [C.c_void_p]; lib.nvbx_stream_kernel.restype = C.c_char_p
        lib.nvbx_stream_destroy.argtypes = [C.c_void_p]
        lib.nvbx_batched_score_host.argtypes = [C.POINTER(AlignerT), C.POINTER(BatchT), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        lib.nvbx_batched_score_host.restype = C.c_int
        lib.nvbx_stream_enact_banded.argtypes = [C.c_void_p, C.POINTER(AlignerT), C.c_int, C.c_int]; lib.nvbx_stream_enact_banded.restype = C.c_int
        lib.nvbx_stream_traceback.argtypes = [C.c_void_p, C.POINTER(AlignerT), C.c_int]; lib.nvbx_stream_traceback.restype = C.c_int
        lib.nvbx_stream_ops_stride.argtypes = [C.c_void_p]; lib.nvbx_stream_ops_stride.restype = C.c_uint32
        lib.nvbx_stream_alignments.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; lib.nvbx_stream_alignments.restype = C.c_int
        lib.nvbx_stream_set_min_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]; lib.nvbx_stream_set_min_scores.restype = C.c_int
        lib.nvbx_set_policy.argtypes = [C.c_int]
        lib.nvbx_kernel_launches.restype = C.c_uint64
        lib.nvbx_last_error.restype = C.c_char_p
        _bound = True
    return lib


def concat(seqs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """list of uint8 arrays -> (concatenated bytes, offsets[n + 1])"""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint32)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    return (np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if seqs else np.zeros(0, np.uint8)), offs


class Batch:
    """Host-side job list (the reference's stream_type strings): keeps the numpy arrays alive behind the POD."""

    def __init__(self, patterns: np.ndarray, pattern_offsets: np.ndarray, texts: np.ndarray, text_offsets: np.ndarray,
                 pair_pattern: Optional[np.ndarray] = None, pair_text: Optional[np.ndarray] = None, n_pairs: Optional[int] = None, alphabet: int = CODES):
        self.patterns = np.ascontiguousarray(patterns, dtype=np.uint8); self.poffs = np.ascontiguousarray(pattern_offsets, dtype=np.uint32)
        self.texts = np.ascontiguousarray(texts, dtype=np.uint8); self.toffs = np.ascontiguousarray(text_offsets, dtype=np.uint32)
        self.pp = None if pair_pattern is None else np.ascontiguousarray(pair_pattern, dtype=np.uint32)
        self.pt = None if pair_text is None else np.ascontiguousarray(pair_text, dtype=np.uint32)
        if n_pairs is None:
            n_pairs = len(self.pp) if self.pp is not None else (len(self.pt) if self.pt is not None else len(self.poffs) - 1)
        self.n_pairs = int(n_pairs)
        self.pod = BatchT(self.patterns.ctypes.data, self.poffs.ctypes.data, len(self.poffs) - 1, self.texts.ctypes.data, self.toffs.ctypes.data,
                          len(self.toffs) - 1, None if self.pp is None else self.pp.ctypes.data, None if self.pt is None else self.pt.ctypes.data,
                          self.n_pairs, alphabet)

    def job(self, k: int) -> Tuple[np.ndarray, np.ndarray]:
        pi = k if self.pp is None else int(self.pp[k])
        ti = (0 if len(self.toffs) == 2 and self.pt is None else k) if self.pt is None else int(self.pt[k])
        return self.patterns[self.poffs[pi]:self.poffs[pi + 1]], self.texts[self.toffs[ti]:self.toffs[ti + 1]]


def aligner(algo: int, type_: int, match: int = 0, mismatch: int = 0, gap_open: int = 0, gap_ext: int = 0) -> AlignerT:
    return AlignerT(algo, type_, match, mismatch, gap_open, gap_ext)


class Stream:
    """Device-resident batch: create (upload + pack) once, enact any number of aligners on it (BatchedAlignmentScore::enact)."""

    def __init__(self, batch: Batch, cuda_stream: Optional[int] = None):
        self.batch = batch
        self._h = _lib().nvbx_stream_create(C.byref(batch.pod), cuda_stream)
        if not self._h:
            raise RuntimeError("nvbx_stream_create: " + _lib().nvbx_last_error().decode())

    def enact(self, al: AlignerT, want_sinks: bool = False) -> None:
        if _lib().nvbx_stream_enact(self._h, C.byref(al), 1 if want_sinks else 0) != 0:
            raise RuntimeError("nvbx_stream_enact: " + _lib().nvbx_last_error().decode())

    def set_min_scores(self, per_job: Optional[np.ndarray] = None, uniform: int = -(1 << 31)) -> None:
        """the reference's per-job context_type::min_score (batched.h:249): early-exit threshold of the following enact calls;
        (None, INT32_MIN) switches it off"""
        arr = None if per_job is None else np.ascontiguousarray(per_job, dtype=np.int32)
        if arr is not None and arr.size != self.batch.n_pairs:
            raise ValueError("set_min_scores: one threshold per job")
        if _lib().nvbx_stream_set_min_scores(self._h, None if arr is None else arr.ctypes.data, int(uniform)) != 0:
            raise RuntimeError("nvbx_stream_set_min_scores: " + _lib().nvbx_last_error().decode())

    def results(self, want_sinks: bool = False):
        n = self.batch.n_pairs
        sc = np.empty(n, np.int32); si = np.empty(n, np.int32) if want_sinks else None; sj = np.empty(n, np.int32) if want_sinks else None
        rc = _lib().nvbx_stream_results(self._h, sc.ctypes.data, None if si is None else si.ctypes.data, None if sj is None else sj.ctypes.data)
        if rc != 0:
            raise RuntimeError("nvbx_stream_results: " + _lib().nvbx_last_error().decode())
        return sc, si, sj

    def enact_banded(self, al: AlignerT, band_len: int, want_sinks: bool = False) -> None:
        """BatchedBandedAlignmentScore<BAND_LEN, stream>::enact (batched.h:337-357)"""
        if _lib().nvbx_stream_enact_banded(self._h, C.byref(al), band_len, 1 if want_sinks else 0) != 0:
            raise RuntimeError("nvbx_stream_enact_banded: " + _lib().nvbx_last_error().decode())

    def traceback(self, al: AlignerT, band_len: int = 0) -> None:
        """BatchedAlignmentTraceback::enact (band_len 0) / BatchedBandedAlignmentTraceback<BAND_LEN>::enact (batched.h:436-478)"""
        if _lib().nvbx_stream_traceback(self._h, C.byref(al), band_len) != 0:
            raise RuntimeError("nvbx_stream_traceback: " + _lib().nvbx_last_error().decode())

    def alignments(self, want_ops: bool = True):
        """(structured array of nvbx_alignment_t, ops[n_pairs, stride] uint8 in push order or None)"""
        n = self.batch.n_pairs
        alns = np.zeros(n, dtype=np.dtype([("score", "<i4"), ("source_text", "<i4"), ("source_pattern", "<i4"), ("sink_text", "<i4"), ("sink_pattern", "<i4"),
                                           ("n_ops", "<u4"), ("clip_end", "<u4"), ("clip_begin", "<u4")]))
        stride = int(_lib().nvbx_stream_ops_stride(self._h))
        ops = np.zeros((n, stride), np.uint8) if want_ops else None
        if _lib().nvbx_stream_alignments(self._h, alns.ctypes.data, None if ops is None else ops.ctypes.data) != 0:
            raise RuntimeError("nvbx_stream_alignments: " + _lib().nvbx_last_error().decode())
        return alns, ops

    def last_ms(self) -> float:
        return float(_lib().nvbx_stream_last_ms(self._h))

    def cells(self) -> int:
        return int(_lib().nvbx_stream_cells(self._h))

    def kernel(self) -> str:
        return _lib().nvbx_stream_kernel(self._h).decode()

    def close(self) -> None:
        if self._h:
            _lib().nvbx_stream_destroy(self._h); self._h = None

    def __enter__(self): return self
    def __exit__(self, *exc): self.close()


def batched_score(al: AlignerT, batch: Batch, want_sinks: bool = False):
    """One-call host-buffer form.  Returns (scores, sink_text, sink_pattern, ms[upload+pack, scoring, download])."""
    n = batch.n_pairs
    sc = np.empty(n, np.int32); si = np.empty(n, np.int32) if want_sinks else None; sj = np.empty(n, np.int32) if want_sinks else None
    ms = (C.c_float * 3)()
    rc = _lib().nvbx_batched_score_host(C.byref(al), C.byref(batch.pod), sc.ctypes.data, None if si is None else si.ctypes.data,
                                        None if sj is None else sj.ctypes.data, ms)
    if rc != 0:
        raise RuntimeError("nvbx_batched_score_host: " + _lib().nvbx_last_error().decode())
    return sc, si, sj, (ms[0], ms[1], ms[2])


def set_policy(p: int) -> None:
    _lib().nvbx_set_policy(p)


def kernel_launches() -> int:
    return int(_lib().nvbx_kernel_launches())


# ---- one process per GPU: jobs are independent, so the path shards with no data-path collective -----------------
def shard_jobs(m_lens: np.ndarray, n_lens: np.ndarray, world: int, block: int = 1024):
    """Per-rank job indices balanced on DP cells: jobs sorted by cells, cut into blocks, blocks dealt by LPT (largest first to the
    least loaded rank).  Deterministic, so every rank can compute its own share without communication."""
    cells = np.asarray(m_lens, np.int64) * np.asarray(n_lens, np.int64)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(world, np.int64)
    per_rank = [[] for _ in range(world)]
    for s in range(0, len(order), block):
        idx = order[s:s + block]
        r = int(np.argmin(load))
        per_rank[r].append(idx); load[r] += int(cells[idx].sum())
    return [np.sort(np.concatenate(x)).astype(np.uint32) if x else np.zeros(0, np.uint32) for x in per_rank]


def sub_batch(batch: Batch, jobs: np.ndarray) -> Batch:
    """The given jobs as a batch of their own (pair index arrays into the SAME strings: nothing is copied or re-packed)."""
    pp = jobs if batch.pp is None else batch.pp[jobs]
    shared = batch.pt is None and len(batch.toffs) == 2
    pt = (np.zeros(len(jobs), np.uint32) if shared else jobs) if batch.pt is None else batch.pt[jobs]
    return Batch(batch.patterns, batch.poffs, batch.texts, batch.toffs, pair_pattern=pp, pair_text=pt, alphabet=batch.pod.alphabet)


def score_sharded(batch: Batch, score_fn, dist=None):
    """Every rank holds the (host) batch, scores its own jobs with `score_fn(sub_batch) -> int32 scores` (on its GPU through the C ABI;
    a CPU stand-in in the gloo test); rank 0 gathers the scores into the caller's job order (None on the other ranks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(score_fn(batch), np.int32)
    rank, world = dist.get_rank(), dist.get_world_size()
    m = np.diff(batch.poffs)[np.arange(batch.n_pairs) if batch.pp is None else batch.pp]
    shared = batch.pt is None and len(batch.toffs) == 2
    n = np.diff(batch.toffs)[(np.zeros(batch.n_pairs, np.int64) if shared else np.arange(batch.n_pairs)) if batch.pt is None else batch.pt]
    mine = shard_jobs(m, n, world)[rank]
    res = np.asarray(score_fn(sub_batch(batch, mine)), np.int32) if len(mine) else np.zeros(0, np.int32)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((mine, res), gathered, dst=0)
    if rank != 0:
        return None
    out = np.zeros(batch.n_pairs, np.int32)
    for idx, r in gathered:
        out[idx] = r
    return out
