"""ctypes front-ends for the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import
this module; the product package (genomics_gpu_b200) never does.

* ``Oracle``  -- oracle/gotoh_compat.c, our plain-C restatement of the reference recurrences.
* ``RefHost`` -- oracle/_ref/libgasal_ref_host_mq<N>.so, the reference's own kernel headers
  (Non-CDP/GASAL2/src/kernels/*.h) compiled unmodified for the host (see oracle/ref_shim/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# enums of the reference (Non-CDP/GASAL2/src/gasal.h:37-72)
WITHOUT_START, WITH_START, WITH_TB = 0, 1, 2
NONE, QUERY, TARGET, BOTH = 0, 1, 2, 3
GLOBAL, SEMI_GLOBAL, LOCAL, MICROLOCAL, BANDED, KSW = 1, 2, 3, 4, 5, 6

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def _ptr(a: Optional[np.ndarray], ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def build_oracle(target: str = "oracle") -> None:
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


class _Scoring(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("match", "mismatch", "gap_open", "gap_extend", "n_code", "n_penalty")]


class _Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("algo", "start_mode", "second_best", "head", "tail", "k_band")] + [("max_query_len", C.c_uint32)]


class _Results(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_int32)) for n in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start",
                                                    "target_batch_start", "aln_score2", "query_batch_end2", "target_batch_end2")] + \
               [("cigar", C.POINTER(C.c_uint8)), ("n_cigar_ops", C.POINTER(C.c_uint32))]


def _alloc_results(n: int, q_bytes: int) -> Dict[str, np.ndarray]:
    r = {k: np.zeros(n, dtype=np.int32) for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start",
                                                 "target_batch_start", "aln_score2", "query_batch_end2", "target_batch_end2")}
    r["cigar"] = np.zeros(max(q_bytes, 8), dtype=np.uint8)
    r["n_cigar_ops"] = np.zeros(n, dtype=np.uint32)
    return r


class Oracle:
    """oracle/gotoh_compat.c"""

    def __init__(self):
        path = os.path.join(HERE, "_build", "libgotoh_compat.so")
        if not os.path.exists(path):
            build_oracle("oracle")
        self.lib = C.CDLL(path)
        self.lib.gc_pack.argtypes = [u8p, C.c_uint32, u32p]
        self.lib.gc_align_batch.restype = C.c_int
        self.lib.gc_revcomp_batch.argtypes = [u32p, u32p, u32p, u32p, u32p, u32p, u8p, u8p, C.c_uint32]

    def pack(self, ascii_bytes: np.ndarray) -> np.ndarray:
        assert ascii_bytes.dtype == np.uint8 and ascii_bytes.size % 8 == 0
        out = np.zeros(ascii_bytes.size // 8, dtype=np.uint32)
        if ascii_bytes.size:
            self.lib.gc_pack(np.ascontiguousarray(ascii_bytes), ascii_bytes.size, out)
        return out

    def revcomp(self, pq, pt, q_lens, t_lens, q_offs, t_offs, q_op, t_op, mode=0):
        """mode 0: padded-array reversal (engine default); 1: the reference as built by nvcc 12.9 for sm_100 (UB artefact)"""
        pq, pt = pq.copy(), pt.copy()
        self.lib.gc_set_revcomp_mode(int(mode))
        self.lib.gc_revcomp_batch(pq, pt, q_lens, t_lens, q_offs, t_offs, np.ascontiguousarray(q_op, dtype=np.uint8),
                                  np.ascontiguousarray(t_op, dtype=np.uint8), len(q_lens))
        return pq, pt

    def align(self, pq, pt, q_lens, t_lens, q_offs, t_offs, *, algo, start=WITHOUT_START, second_best=False, head=TARGET,
              tail=TARGET, k_band=0, scoring=(1, 4, 6, 1), seed_scores=None, max_query_len=0, n_threads=0,
              n_code=0x4E, n_penalty=-1) -> Dict[str, np.ndarray]:
        n = len(q_lens)
        res = _alloc_results(n, pq.size * 8)
        sc = _Scoring(*scoring, n_code, n_penalty)
        p = _Params(algo, start, int(bool(second_best)), head, tail, k_band, max_query_len)
        rs = _Results(*[_ptr(res[k], C.c_int32) for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start",
                                                          "target_batch_start", "aln_score2", "query_batch_end2", "target_batch_end2")],
                      _ptr(res["cigar"], C.c_uint8), _ptr(res["n_cigar_ops"], C.c_uint32))
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (pq, pt, q_lens, t_lens, q_offs, t_offs)]
        seed = None if seed_scores is None else np.ascontiguousarray(seed_scores, dtype=np.uint32)
        rc = self.lib.gc_align_batch(C.byref(sc), C.byref(p), *[_ptr(a, C.c_uint32) for a in arrs], _ptr(seed, C.c_uint32),
                                     C.c_uint32(n), C.byref(rs), C.c_int(n_threads))
        if rc != 0:
            raise RuntimeError("gc_align_batch failed")
        return res


class RefHost:
    """The reference's own kernels on the host (oracle/_ref/libgasal_ref_host_mq<N>.so)."""

    def __init__(self, max_query_len: int = 1024):
        path = os.path.join(HERE, "_ref", f"libgasal_ref_host_mq{max_query_len}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref_host` where /root/reference exists)")
        self.lib = C.CDLL(path)
        self.max_query_len = self.lib.ref_max_query_len()
        self.lib.ref_align.restype = C.c_int

    @staticmethod
    def available(max_query_len: int = 1024) -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", f"libgasal_ref_host_mq{max_query_len}.so"))

    def pack(self, uq: np.ndarray, ut: np.ndarray, n_alns: int):
        pq = np.zeros(uq.size // 8, dtype=np.uint32)
        pt = np.zeros(ut.size // 8, dtype=np.uint32)
        self.lib.ref_pack(_ptr(uq, C.c_uint8), _ptr(ut, C.c_uint8), _ptr(pq, C.c_uint32), _ptr(pt, C.c_uint32),
                          C.c_uint32(uq.size), C.c_uint32(ut.size), C.c_uint32(n_alns))
        return pq, pt

    def align(self, pq, pt, q_lens, t_lens, q_offs, t_offs, *, algo, start=WITHOUT_START, second_best=False, head=TARGET,
              tail=TARGET, k_band=0, scoring=(1, 4, 6, 1), seed_scores=None, n_threads=0) -> Dict[str, np.ndarray]:
        n = len(q_lens)
        assert int(max(q_lens)) <= self.max_query_len
        res = _alloc_results(n, pq.size * 8)
        self.lib.ref_set_scores(*[C.c_int(int(v)) for v in scoring])
        arrs = [np.ascontiguousarray(a, dtype=np.uint32).copy() for a in (pq, pt, q_lens, t_lens, q_offs, t_offs)]
        seed = None if seed_scores is None else np.ascontiguousarray(seed_scores, dtype=np.uint32)
        rc = self.lib.ref_align(algo, start, int(bool(second_best)), head, tail, k_band, *[_ptr(a, C.c_uint32) for a in arrs],
                                _ptr(seed, C.c_uint32), n,
                                *[_ptr(res[k], C.c_int32) for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start",
                                                                    "target_batch_start", "aln_score2", "query_batch_end2", "target_batch_end2")],
                                _ptr(res["cigar"], C.c_uint8), _ptr(res["n_cigar_ops"], C.c_uint32), n_threads)
        if rc != 0:
            raise RuntimeError("ref_align failed")
        return res


class RefGpu:
    """The UNMODIFIED reference library built for sm_100 (oracle/_ref/ref_gpu_driver_mq<N>, see
    oracle/ref_gpu/), run as a subprocess on a batch file.  Needs a GPU.  Used (a) to pin the oracle
    against the reference's real GPU build, including gasal_reversecomplement_kernel whose C is only
    well-defined on the GPU, and (b) as the same-GPU baseline bench.py reports."""

    def __init__(self, max_query_len: int = 160):
        self.path = os.path.join(HERE, "_ref", f"ref_gpu_driver_mq{max_query_len}")
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.max_query_len = max_query_len

    @staticmethod
    def available(max_query_len: int = 160) -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", f"ref_gpu_driver_mq{max_query_len}"))

    def run(self, batch, *, algo, start=WITHOUT_START, second_best=False, head=TARGET, tail=TARGET, k_band=0, scoring=(1, 4, 6, 1),
            q_op=None, t_op=None, batch_size=5000, reps=1, workdir=None):
        import json
        import tempfile
        n = batch.n
        q_op = np.zeros(n, np.uint8) if q_op is None else np.ascontiguousarray(q_op, dtype=np.uint8)
        t_op = np.zeros(n, np.uint8) if t_op is None else np.ascontiguousarray(t_op, dtype=np.uint8)
        revcomp = int(bool(q_op.any() or t_op.any()))
        with tempfile.TemporaryDirectory(dir=workdir) as d:
            fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
            with open(fin, "wb") as f:
                np.array([n, batch.q_bytes.size, batch.t_bytes.size], dtype=np.uint32).tofile(f)
                for a in (batch.q_bytes, batch.t_bytes, batch.q_offs, batch.t_offs, batch.q_lens, batch.t_lens, q_op, t_op):
                    np.ascontiguousarray(a).tofile(f)
            cmd = [self.path, fin, fout] + [str(int(x)) for x in (algo, start, int(bool(second_best)), head, tail, k_band, revcomp, *scoring, batch_size, reps)]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"reference GPU driver failed: {p.stderr[-2000:]}")
            timing = json.loads(p.stdout.strip().splitlines()[-1])
            raw = np.fromfile(fout, dtype=np.uint8)
        nn = int(raw[:4].view(np.uint32)[0])
        body = raw[4:4 + 4 * 9 * nn].view(np.int32).reshape(9, nn)
        keys = ["aln_score", "query_batch_end", "target_batch_end", "query_batch_start", "target_batch_start", "aln_score2", "query_batch_end2", "target_batch_end2"]
        res = {k: body[i].copy() for i, k in enumerate(keys)}
        res["n_cigar_ops"] = body[8].view(np.uint32).copy()
        off = 4 + 4 * 9 * nn
        qb = int(raw[off:off + 4].view(np.uint32)[0])
        res["cigar"] = raw[off + 4:off + 4 + qb].copy()
        res["_timing"] = timing
        return res


# ---- full-matrix linear-gap NW (oracle/nw_linear.c; reference Non-CDP/NW/main.cu:9-27) -------------------------------
class NwOracle:
    """oracle/nw_linear.c: nw_linear_matrix / nw_linear_checksum."""

    def __init__(self):
        path = os.path.join(HERE, "_build", "libnw_linear.so")
        if not os.path.exists(path):
            build_oracle("oracle")
        self.lib = C.CDLL(path)
        self.lib.nw_linear_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int]
        self.lib.nw_linear_checksum.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int]
        self.lib.nw_linear_checksum.restype = C.c_uint64

    def matrix(self, reference: np.ndarray, query: np.ndarray, scores=(1, -1, -1, -1)) -> np.ndarray:
        reference = np.ascontiguousarray(reference, dtype=np.uint8); query = np.ascontiguousarray(query, dtype=np.uint8)
        n = reference.size
        out = np.empty((n, n), dtype=np.int32)
        self.lib.nw_linear_matrix(reference.ctypes.data, query.ctypes.data, out.ctypes.data, n, *scores)
        return out

    def checksum(self, reference: np.ndarray, query: np.ndarray, scores=(1, -1, -1, -1)) -> int:
        reference = np.ascontiguousarray(reference, dtype=np.uint8); query = np.ascontiguousarray(query, dtype=np.uint8)
        return int(self.lib.nw_linear_checksum(reference.ctypes.data, query.ctypes.data, reference.size, *scores))


def nw_checksum_of(matrix: np.ndarray) -> int:
    """The hash of include/nw_b200.h (nwx_checksum) over a host matrix."""
    flat = matrix.reshape(-1).astype(np.uint32).astype(np.uint64)
    idx = np.arange(flat.size, dtype=np.uint64)
    w = ((idx * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) | np.uint64(1)
    return int((flat * w).sum(dtype=np.uint64))


class NwRefHost:
    """oracle/_ref/libnw_ref_host.so: the reference's own nw_cpu and generateQuery (NW/main.cu) compiled for the host."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libnw_ref_host.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.nw_ref_cpu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
        self.lib.nw_ref_generate_query.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint]

    def matrix(self, reference: np.ndarray, query: np.ndarray) -> np.ndarray:
        reference = np.ascontiguousarray(reference, dtype=np.uint8); query = np.ascontiguousarray(query, dtype=np.uint8)
        n = reference.size
        out = np.empty((n, n), dtype=np.int32)
        self.lib.nw_ref_cpu(reference.ctypes.data, query.ctypes.data, out.ctypes.data, n)
        return out

    def generate(self, n: int, seed: int):
        reference = np.empty(n, dtype=np.uint8); query = np.empty(n, dtype=np.uint8)
        self.lib.nw_ref_generate_query(reference.ctypes.data, query.ctypes.data, n, seed)
        return reference, query


# ---- nvbio-style batched scoring (oracle/nvbio_compat.c; reference Non-CDP/NvB/nvbio/alignment/gotoh/gotoh_inl.h:473-900) ----
NVB_GLOBAL, NVB_LOCAL, NVB_SEMI_GLOBAL = 0, 1, 2
NVB_GOTOH, NVB_SW, NVB_ED = 0, 1, 2
NVB_BAND = {NVB_GOTOH: 8, NVB_SW: 16, NVB_ED: 16}     # the reference's BAND_LEN per aligner (gotoh_inl.h:1494-1503, sw_inl.h:1333-1340)


class NvbioOracle:
    """oracle/nvbio_compat.c (restatement) and, when built, oracle/_ref/libnvbio_ref.so (the reference's own host code path)."""

    def __init__(self):
        path = os.path.join(HERE, "_build", "libnvbio_compat.so")
        if not os.path.exists(path):
            build_oracle("oracle")
        self.lib = C.CDLL(path)
        self.lib.nvbio_compat_score.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        self.lib.nvbio_compat_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                                C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        ref = os.path.join(HERE, "_ref", "libnvbio_ref.so")
        self.ref = C.CDLL(ref) if os.path.exists(ref) else None
        V, U = C.c_void_p, C.c_uint32
        self.lib.nvbio_compat_score_min.argtypes = [C.c_int, C.c_int, V, V, U, V, U, U, C.c_int32, V]
        self.lib.nvbio_compat_batch_min.argtypes = [C.c_int, C.c_int, V, V, V, V, V, U, V, V, U, U, V, C.c_int32, V, V, V]
        if self.ref is not None and hasattr(self.ref, "nvbio_ref_score_min"):
            self.ref.nvbio_ref_score_min.argtypes = [C.c_int, C.c_int, V, V, U, V, U, C.c_int32, V]
        self.lib.nvbio_compat_banded_score.argtypes = [C.c_int, C.c_int, C.c_int, V, V, U, V, U, V]
        self.lib.nvbio_compat_traceback.argtypes = [C.c_int, C.c_int, V, V, U, V, U, U, V, V, U]
        self.lib.nvbio_compat_banded_traceback.argtypes = [C.c_int, C.c_int, C.c_int, V, V, U, V, U, V, V, U]
        if self.ref is not None:
            self.ref.nvbio_ref_score.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
            if hasattr(self.ref, "nvbio_ref_traceback"):
                self.ref.nvbio_ref_banded_score.argtypes = [C.c_int, C.c_int, C.c_int, V, V, U, V, U, V]
                self.ref.nvbio_ref_traceback.argtypes = [C.c_int, C.c_int, V, V, U, V, U, V, V, U]
                self.ref.nvbio_ref_banded_traceback.argtypes = [C.c_int, C.c_int, C.c_int, V, V, U, V, U, V, V, U]

    @staticmethod
    def _scheme(scheme):
        return np.asarray(list(scheme) + [0] * (4 - len(scheme)), dtype=np.int32)

    def score(self, algo, type_, scheme, pattern, text):
        p = np.ascontiguousarray(pattern, dtype=np.uint8); t = np.ascontiguousarray(text, dtype=np.uint8)
        sc = self._scheme(scheme); out = np.zeros(3, np.int32)
        rc = self.lib.nvbio_compat_score(algo, type_, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, NVB_BAND[algo], out.ctypes.data)
        assert rc == 0
        return tuple(int(x) for x in out)

    def ref_score(self, algo, type_, scheme, pattern, text):
        p = np.ascontiguousarray(pattern, dtype=np.uint8); t = np.ascontiguousarray(text, dtype=np.uint8)
        sc = self._scheme(scheme); out = np.zeros(3, np.int32)
        rc = self.ref.nvbio_ref_score(algo, type_, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, out.ctypes.data)
        assert rc == 0
        return tuple(int(x) for x in out)

    def score_min(self, algo, type_, scheme, pattern, text, min_score, ref=False):
        """alignment_score with a min_score threshold: (score, sink text end, sink pattern end, ok); ok = 0 when it gave up early"""
        p = np.ascontiguousarray(pattern, dtype=np.uint8); t = np.ascontiguousarray(text, dtype=np.uint8)
        sc = self._scheme(scheme); out = np.zeros(4, np.int32)
        if ref:
            rc = self.ref.nvbio_ref_score_min(algo, type_, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, int(min_score), out.ctypes.data)
        else:
            rc = self.lib.nvbio_compat_score_min(algo, type_, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, NVB_BAND[algo], int(min_score),
                                                 out.ctypes.data)
        assert rc == 0, rc
        return tuple(int(x) for x in out)

    def batch_min(self, algo, type_, scheme, patterns, poffs, texts, toffs, min_score=-(1 << 31), min_scores=None):
        patterns = np.ascontiguousarray(patterns, dtype=np.uint8); texts = np.ascontiguousarray(texts, dtype=np.uint8)
        poffs = np.ascontiguousarray(poffs, dtype=np.uint32); toffs = np.ascontiguousarray(toffs, dtype=np.uint32)
        n_pairs = len(poffs) - 1
        ms = None if min_scores is None else np.ascontiguousarray(min_scores, dtype=np.int32)
        sc = self._scheme(scheme)
        s = np.empty(n_pairs, np.int32); si = np.empty(n_pairs, np.int32); sj = np.empty(n_pairs, np.int32)
        rc = self.lib.nvbio_compat_batch_min(algo, type_, sc.ctypes.data, patterns.ctypes.data, poffs.ctypes.data, texts.ctypes.data, toffs.ctypes.data,
                                             len(toffs) - 1, None, None, n_pairs, NVB_BAND[algo], None if ms is None else ms.ctypes.data, int(min_score),
                                             s.ctypes.data, si.ctypes.data, sj.ctypes.data)
        assert rc == 0
        return s, si, sj

    # ---- banded scoring and traceback (banded_inl.h, alignment_inl.h:361-467); ref=True runs the reference's own host code ----
    def banded_score(self, algo, type_, band, scheme, pattern, text, ref=False):
        """(score, sink text end, sink pattern end, ok)"""
        p = np.ascontiguousarray(pattern, dtype=np.uint8); t = np.ascontiguousarray(text, dtype=np.uint8)
        sc = self._scheme(scheme); out = np.zeros(4, np.int32)
        fn = self.ref.nvbio_ref_banded_score if ref else self.lib.nvbio_compat_banded_score
        rc = fn(algo, type_, band, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, out.ctypes.data)
        assert rc == 0, rc
        return tuple(int(x) for x in out)

    def traceback(self, algo, type_, scheme, pattern, text, band=0, ref=False):
        """band = 0: full-matrix alignment_traceback; else banded_alignment_traceback<band>.  Returns a dict: score, source (text,
        pattern), sink (text, pattern), ops (uint8, push order: from the END of the alignment to its source; 0 substitution,
        1 insertion, 2 deletion), clip_end, clip_begin."""
        p = np.ascontiguousarray(pattern, dtype=np.uint8); t = np.ascontiguousarray(text, dtype=np.uint8)
        sc = self._scheme(scheme); out = np.zeros(9, np.int32); cap = p.size + t.size + 8; ops = np.zeros(cap, np.uint8)
        if band:
            fn = self.ref.nvbio_ref_banded_traceback if ref else self.lib.nvbio_compat_banded_traceback
            rc = fn(algo, type_, band, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, out.ctypes.data, ops.ctypes.data, cap)
        elif ref:
            rc = self.ref.nvbio_ref_traceback(algo, type_, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, out.ctypes.data, ops.ctypes.data, cap)
        else:
            rc = self.lib.nvbio_compat_traceback(algo, type_, sc.ctypes.data, p.ctypes.data, p.size, t.ctypes.data, t.size, NVB_BAND[algo],
                                                 out.ctypes.data, ops.ctypes.data, cap)
        assert rc == 0, rc
        return dict(score=int(out[0]), source=(int(out[1]), int(out[2])), sink=(int(out[3]), int(out[4])), ops=ops[:int(out[5])].copy(),
                    clip_end=int(out[6]), clip_begin=int(out[7]))

    def batch(self, algo, type_, scheme, patterns, poffs, texts, toffs, pair_pattern=None, pair_text=None, n_pairs=None):
        patterns = np.ascontiguousarray(patterns, dtype=np.uint8); texts = np.ascontiguousarray(texts, dtype=np.uint8)
        poffs = np.ascontiguousarray(poffs, dtype=np.uint32); toffs = np.ascontiguousarray(toffs, dtype=np.uint32)
        pp = None if pair_pattern is None else np.ascontiguousarray(pair_pattern, dtype=np.uint32)
        pt = None if pair_text is None else np.ascontiguousarray(pair_text, dtype=np.uint32)
        if n_pairs is None:
            n_pairs = len(pp) if pp is not None else (len(pt) if pt is not None else len(poffs) - 1)
        sc = self._scheme(scheme)
        s = np.empty(n_pairs, np.int32); si = np.empty(n_pairs, np.int32); sj = np.empty(n_pairs, np.int32)
        rc = self.lib.nvbio_compat_batch(algo, type_, sc.ctypes.data, patterns.ctypes.data, poffs.ctypes.data, texts.ctypes.data, toffs.ctypes.data,
                                         len(toffs) - 1, None if pp is None else pp.ctypes.data, None if pt is None else pt.ctypes.data, n_pairs,
                                         NVB_BAND[algo], s.ctypes.data, si.ctypes.data, sj.ctypes.data)
        assert rc == 0
        return s, si, sj


def nvb_rle(ops, order="push") -> str:
    """run-length string of a traceback's ops in the letters of the reference's test program (M / I / D,
    NvB/nvbio-test/alignment_test.cu); order="push" keeps the reference's push order (end of the alignment first)"""
    seq = list(ops) if order == "push" else list(ops)[::-1]
    out, i = [], 0
    while i < len(seq):
        j = i
        while j < len(seq) and seq[j] == seq[i]:
            j += 1
        out.append(f"{j - i}{'MID'[seq[i]]}")
        i = j
    return "".join(out)
