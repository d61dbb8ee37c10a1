"""Host-side driver over the C ABI, mirroring how the reference's only caller uses GASAL2
(Non-CDP/GASAL2/test_prog/test_prog.cpp:203-437): init storages -> fill host batches -> gasal_aln_async ->
poll gasal_is_aln_async_done -> read host_res.  All compute happens in libgasal_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, Optional, Sequence

import numpy as np

from . import capi
from .synth import PairBatch


def _np_view(ptr, n, dtype):
    if not ptr:
        return None
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


class Aligner:
    """One thread's worth of GASAL2 state: `n_streams` storages (reference NB_STREAMS, test_prog.cpp:12)."""

    def __init__(self, *, algo: int, start: int = capi.WITHOUT_START, second_best: bool = False, head: int = capi.TARGET,
                 tail: int = capi.TARGET, k_band: int = 0, is_packed: bool = False, reverse_complement: bool = False,
                 scoring: Sequence[int] = (1, 4, 6, 1), max_query_len: int = 160, max_target_len: int = 320,
                 max_n_alns: int = 5000, n_streams: int = 2, device: Optional[int] = None):
        self.lib = capi.load()
        if device is not None:
            self.lib.gasal_set_device(device, 0)
        self.params = capi.Params()
        self.lib.gasal_params_default(C.byref(self.params))
        self.params.algo = algo; self.params.start_pos = start; self.params.secondBest = int(bool(second_best))
        self.params.semiglobal_skipping_head = head; self.params.semiglobal_skipping_tail = tail; self.params.k_band = k_band
        self.params.isPacked = int(is_packed); self.params.isReverseComplement = int(reverse_complement)
        sc = capi.SubstScores(*[int(v) for v in scoring])
        self.lib.gasal_copy_subst_scores(C.byref(sc))
        self.vec = self.lib.gasal_init_gpu_storage_v(n_streams)
        self.lib.gasal_init_streams(C.byref(self.vec), max_query_len, max_target_len, max_n_alns, C.byref(self.params))
        self.max_n_alns = max_n_alns
        self._closed = False

    def storage(self, i: int) -> capi.GpuStorage:
        return self.vec.a[i]

    # -- filling: test_prog.cpp:279-315 ----------------------------------------------------------------
    def fill(self, i: int, batch: PairBatch, q_op: Optional[np.ndarray] = None, t_op: Optional[np.ndarray] = None) -> None:
        """Stage a whole PairBatch (already in gasal_host_batch_fill layout) into storage i with two bulk
        gasal_host_batch_add calls, then write offsets / lens / ops as the reference's caller does."""
        s = self.vec.a[i]
        n = batch.n
        if n > s.host_max_n_alns:
            self.lib.gasal_host_alns_resize(C.byref(s), max(n, 2 * s.host_max_n_alns), C.byref(self.params))
        self.lib.gasal_host_batch_add(C.byref(s), 0, batch.q_bytes.ctypes.data_as(C.c_char_p), batch.q_bytes.size, capi.QUERY)
        self.lib.gasal_host_batch_add(C.byref(s), 0, batch.t_bytes.ctypes.data_as(C.c_char_p), batch.t_bytes.size, capi.TARGET)
        C.memmove(s.host_query_batch_offsets, batch.q_offs.ctypes.data, 4 * n)
        C.memmove(s.host_target_batch_offsets, batch.t_offs.ctypes.data, 4 * n)
        C.memmove(s.host_query_batch_lens, batch.q_lens.ctypes.data, 4 * n)
        C.memmove(s.host_target_batch_lens, batch.t_lens.ctypes.data, 4 * n)
        if q_op is not None:
            q_op = np.ascontiguousarray(q_op, dtype=np.uint8); t_op = np.ascontiguousarray(t_op, dtype=np.uint8)
            self.lib.gasal_op_fill(C.byref(s), q_op.ctypes.data_as(C.POINTER(C.c_uint8)), n, capi.QUERY)
            self.lib.gasal_op_fill(C.byref(s), t_op.ctypes.data_as(C.POINTER(C.c_uint8)), n, capi.TARGET)
        s.current_n_alns = n

    def fill_one_by_one(self, i: int, queries: Sequence[bytes], targets: Sequence[bytes]) -> None:
        """The literal per-sequence loop of test_prog.cpp:289-313 (gasal_host_batch_fill pads with 'N')."""
        s = self.vec.a[i]
        qi = ti = 0
        for j, (q, t) in enumerate(zip(queries, targets)):
            s.current_n_alns += 1
            if s.current_n_alns > s.host_max_n_alns:
                self.lib.gasal_host_alns_resize(C.byref(s), s.host_max_n_alns * 2, C.byref(self.params))
            s.host_query_batch_offsets[j] = qi; s.host_target_batch_offsets[j] = ti
            qi = self.lib.gasal_host_batch_fill(C.byref(s), qi, q, len(q), capi.QUERY)
            ti = self.lib.gasal_host_batch_fill(C.byref(s), ti, t, len(t), capi.TARGET)
            s.host_query_batch_lens[j] = len(q); s.host_target_batch_lens[j] = len(t)
        self._last_bytes = (qi, ti)

    def set_seed_scores(self, i: int, seeds: np.ndarray) -> None:
        s = self.vec.a[i]
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        C.memmove(s.host_seed_scores, seeds.ctypes.data, 4 * seeds.size)

    def launch(self, i: int, q_bytes: int, t_bytes: int, n: int) -> None:
        self.lib.gasal_aln_async(C.byref(self.vec.a[i]), q_bytes, t_bytes, n, C.byref(self.params))

    def done(self, i: int) -> int:
        return self.lib.gasal_is_aln_async_done(C.byref(self.vec.a[i]))

    def wait(self, i: int) -> None:
        while True:
            r = self.done(i)
            if r == 0:
                return
            if r == -2:
                raise RuntimeError("gasal_is_aln_async_done: nothing was launched on this storage")
            time.sleep(0)

    def results(self, i: int, n: int, q_bytes: int = 0) -> Dict[str, np.ndarray]:
        """Copies of host_res (valid after done() returned 0)."""
        s = self.vec.a[i]
        out: Dict[str, np.ndarray] = {}
        r = s.host_res.contents
        for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start", "target_batch_start"):
            v = _np_view(getattr(r, k), n, np.int32)
            if v is not None:
                out[k] = v.copy()
        if self.params.start_pos == capi.WITH_TB and r.cigar and r.n_cigar_ops:
            out["cigar"] = _np_view(r.cigar, q_bytes, np.uint8).copy()
            out["n_cigar_ops"] = _np_view(r.n_cigar_ops, n, np.uint32).copy()
        if self.params.secondBest and s.host_res_second:
            r2 = s.host_res_second.contents
            for k in ("aln_score", "query_batch_end", "target_batch_end"):
                v = _np_view(getattr(r2, k), n, np.int32)
                if v is not None:
                    out[k + "2"] = v.copy()
        return out

    def align(self, batch: PairBatch, i: int = 0, q_op=None, t_op=None, seeds=None, cigar_text_style: Optional[int] = None,
              cigar_stride: int = 64) -> Dict[str, np.ndarray]:
        """fill -> gasal_aln_async -> wait -> results, for one batch on storage i.  With cigar_text_style (0 reference
        letters / 1 SAM) the CIGAR text is formatted on the device in the same stream: out["cigar_text"] is a list of str."""
        self.fill(i, batch, q_op, t_op)
        if seeds is not None:
            self.set_seed_scores(i, seeds)
        self.launch(i, batch.q_bytes.size, batch.t_bytes.size, batch.n)
        text = lens = None
        if cigar_text_style is not None:
            text = np.zeros(batch.n * cigar_stride, dtype=np.uint8); lens = np.zeros(batch.n, dtype=np.uint32)
            rc = self.lib.gasalx_cigar_text_async(C.byref(self.vec.a[i]), cigar_text_style, text.ctypes.data, cigar_stride, lens.ctypes.data)
            if rc != 0:
                raise RuntimeError("gasalx_cigar_text_async: the storage holds no traceback results (start must be WITH_TB)")
        self.wait(i)
        out = self.results(i, batch.n, batch.q_bytes.size)
        if text is not None:
            out["cigar_text_lens"] = lens
            rows = text.reshape(batch.n, cigar_stride)
            out["cigar_text"] = [rows[p, :min(int(lens[p]) & 0x7FFFFFFF, cigar_stride - 1)].tobytes().decode() for p in range(batch.n)]
        return out

    def last_ms(self, i: int):
        buf = (C.c_float * 3)()
        rc = self.lib.gasalx_last_batch_ms(C.byref(self.vec.a[i]), buf)
        return None if rc != 0 else (buf[0], buf[1], buf[2])

    def kernel_name(self, i: int) -> str:
        return self.lib.gasalx_last_kernel_name(C.byref(self.vec.a[i])).decode()

    def launches(self, i: int) -> int:
        return int(self.lib.gasalx_kernel_launches(C.byref(self.vec.a[i])))

    def close(self) -> None:
        if not self._closed:
            self.lib.gasal_destroy_streams(C.byref(self.vec), C.byref(self.params))
            self.lib.gasal_destroy_gpu_storage_v(C.byref(self.vec))
            self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def plan_batches(q_lens: np.ndarray, t_lens: np.ndarray, *, n_gpus: int = 1, max_batch_pairs: int = 262144, max_batch_bytes: int = 1 << 30,
                 max_batch_cells: int = 1 << 36, keep_input_order: bool = False):
    """Pure host planning of the batcher (no GPU needed): length-binned batches and their LPT assignment to GPUs."""
    lib = capi.load()
    sched = capi.Sched(); lib.gasalx_sched_default(C.byref(sched))
    sched.n_gpus = n_gpus; sched.max_batch_pairs = max_batch_pairs; sched.max_batch_bytes = max_batch_bytes
    sched.max_batch_cells = max_batch_cells; sched.keep_input_order = int(keep_input_order)
    ql = np.ascontiguousarray(q_lens, dtype=np.uint32); tl = np.ascontiguousarray(t_lens, dtype=np.uint32)
    n = ql.size
    nb = int(lib.gasalx_plan_batches(ql.ctypes.data, tl.ctypes.data, n, C.byref(sched), None, None, None, None, 0))
    order = np.zeros(n, dtype=np.uint32); start = np.zeros(nb + 1, dtype=np.uint32); gpu = np.zeros(nb, dtype=np.int32); cells = np.zeros(nb, dtype=np.uint64)
    lib.gasalx_plan_batches(ql.ctypes.data, tl.ctypes.data, n, C.byref(sched), order.ctypes.data, start.ctypes.data, gpu.ctypes.data, cells.ctypes.data, nb)
    return {"order": order, "batch_start": start, "batch_gpu": gpu, "batch_cells": cells}


def align_all(batch: PairBatch, *, algo: int, start: int = capi.WITHOUT_START, second_best: bool = False, head: int = capi.TARGET,
              tail: int = capi.TARGET, k_band: int = 0, scoring: Sequence[int] = (1, 4, 6, 1), n_gpus: int = 1, gpu_ids: Optional[Sequence[int]] = None,
              streams_per_gpu: int = 0, max_batch_pairs: int = 0, max_batch_cells: int = 1 << 36, q_op=None, t_op=None,
              keep_input_order: bool = False, host_threads_per_gpu: int = 0, out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
    """The whole job through the library-side batcher (gasalx_align_all): length binning, LPT sharding over the GPUs of this
    process, pinned streaming on per-storage streams, host-side gather in the caller's order.  `out`: the dict of a previous call
    on a job of the same size, to write the results into the same (already touched) arrays."""
    lib = capi.load()
    params = capi.Params(); lib.gasal_params_default(C.byref(params))
    params.algo = algo; params.start_pos = start; params.secondBest = int(bool(second_best))
    params.semiglobal_skipping_head = head; params.semiglobal_skipping_tail = tail; params.k_band = k_band
    params.isReverseComplement = int(q_op is not None or t_op is not None)
    sc = capi.SubstScores(*[int(v) for v in scoring]); lib.gasal_copy_subst_scores(C.byref(sc))
    sched = capi.Sched(); lib.gasalx_sched_default(C.byref(sched))
    sched.n_gpus = n_gpus; sched.streams_per_gpu = streams_per_gpu; sched.max_batch_pairs = max_batch_pairs; sched.max_batch_cells = max_batch_cells
    sched.keep_input_order = int(keep_input_order); sched.host_threads_per_gpu = int(host_threads_per_gpu)
    ids = None
    if gpu_ids is not None:
        ids = (C.c_int32 * n_gpus)(*gpu_ids); sched.gpu_ids = ids
    n = batch.n
    if out is None or len(out.get("aln_score", ())) != n:
        out = {k: np.zeros(n, dtype=np.int32) for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start", "target_batch_start",
                                                        "aln_score2", "query_batch_end2", "target_batch_end2")}
        out["n_cigar_ops"] = np.zeros(n, dtype=np.uint32)
        out["cigar"] = np.zeros(batch.q_bytes.size if start == capi.WITH_TB else 8, dtype=np.uint8)
    res = capi.Results(*[out[k].ctypes.data_as(C.POINTER(C.c_int32)) for k in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start",
                                                                                "target_batch_start", "aln_score2", "query_batch_end2", "target_batch_end2")],
                       out["cigar"].ctypes.data_as(C.POINTER(C.c_uint8)), out["n_cigar_ops"].ctypes.data_as(C.POINTER(C.c_uint32)))
    qo = np.ascontiguousarray(batch.q_offs, dtype=np.uint64); to = np.ascontiguousarray(batch.t_offs, dtype=np.uint64)
    qop = None if q_op is None else np.ascontiguousarray(q_op, dtype=np.uint8)
    top = None if t_op is None else np.ascontiguousarray(t_op, dtype=np.uint8)
    gpu_ms = np.zeros(n_gpus, dtype=np.float64)
    rc = lib.gasalx_align_all(batch.q_bytes.ctypes.data, batch.t_bytes.ctypes.data, qo.ctypes.data, to.ctypes.data, batch.q_lens.ctypes.data,
                              batch.t_lens.ctypes.data, None if qop is None else qop.ctypes.data, None if top is None else top.ctypes.data, None,
                              n, C.byref(params), C.byref(sched), C.byref(res), gpu_ms.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"gasalx_align_all failed ({rc})")
    out["_gpu_ms"] = gpu_ms
    return out


def read_fasta_pairs(query_path: str, target_path: str, n_code: int = 0x4E):
    """Two FASTA files (plain or gzip) -> (PairBatch, q_op, t_op, q_names, t_names) through the native reader
    (gasalx_read_fasta_pairs; header conventions of test_prog.cpp:79-89).  Offsets are 64-bit; a PairBatch is what
    align_all() / Aligner.fill() take (slice it to keep one batch within the 32-bit offsets of the GASAL2 API)."""
    lib = capi.load()
    fp = capi.FastaPairs()
    rc = lib.gasalx_read_fasta_pairs(query_path.encode(), target_path.encode(), n_code, C.byref(fp))
    if rc != 0:
        raise RuntimeError(f"gasalx_read_fasta_pairs failed ({rc})")
    try:
        n = int(fp.n)
        arr = lambda ptr, cnt, dt: np.ctypeslib.as_array(ptr, shape=(max(cnt, 1),))[:cnt].view(dt).copy() if cnt else np.zeros(0, dtype=dt)
        batch = PairBatch(arr(fp.q, int(fp.q_bytes), np.uint8), arr(fp.t, int(fp.t_bytes), np.uint8), arr(fp.q_offs, n, np.uint64), arr(fp.t_offs, n, np.uint64),
                          arr(fp.q_lens, n, np.uint32), arr(fp.t_lens, n, np.uint32))
        q_op, t_op = arr(fp.q_op, n, np.uint8), arr(fp.t_op, n, np.uint8)
        qn = C.string_at(fp.q_names).decode().split("\n")[:n]; tn = C.string_at(fp.t_names).decode().split("\n")[:n]
    finally:
        lib.gasalx_free_fasta_pairs(C.byref(fp))
    return batch, q_op, t_op, qn, tn


def pack_host(ascii_bytes: np.ndarray) -> np.ndarray:
    """4-bit packing on the host in the device layout (for isPacked storages)."""
    a = np.ascontiguousarray(ascii_bytes, dtype=np.uint8)
    assert a.size % 8 == 0
    out = np.zeros(a.size // 8, dtype=np.uint32)
    capi.load().gasalx_pack_host(a.ctypes.data, a.size, out.ctypes.data)
    return out


def cigar_string(cigar_bytes: np.ndarray, n_ops: int) -> str:
    """gasal_res_t::cigar bytes of one pair (reverse order) -> '12M1X3D...' as the reference's test_prog prints it."""
    c = np.ascontiguousarray(cigar_bytes[:n_ops], dtype=np.uint8)
    buf = C.create_string_buffer(16 * max(n_ops, 1) + 16)
    n = capi.load().gasalx_cigar_string(c.ctypes.data, n_ops, buf, len(buf))
    if n < 0:
        raise RuntimeError("cigar buffer too small")
    return buf.value.decode()


def sam_record(qname: str, flag: int, rname: str, pos1: int, mapq: int, cigar: str, seq: Optional[bytes], qual: Optional[bytes],
               score: int, score2: int = -1) -> str:
    """One SAM line through the C ABI (gasalx_sam_record)."""
    n = len(seq) if seq else 0
    buf = C.create_string_buffer(256 + len(qname) + len(rname) + len(cigar) + 2 * n)
    r = capi.load().gasalx_sam_record(buf, len(buf), qname.encode(), flag, rname.encode(), pos1, mapq, cigar.encode(), seq, n, qual, score, score2)
    if r < 0:
        raise RuntimeError("sam buffer too small")
    return buf.value.decode()
