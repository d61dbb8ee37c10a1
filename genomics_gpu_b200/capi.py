"""ctypes binding of the engine's C ABI (include/gasal_b200.h).

This is the binding a Python caller of GASAL2 would write against libgasal_b200.so; it mirrors the
structs of the header field by field.  There is no fallback: if the shared library (built in-tree by
``__graft_entry__.build()`` / ``make -C genomics_gpu_b200/csrc``) is missing, importing raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GASAL_B200_LIB") or os.path.join(HERE, "lib", "libgasal_b200.so")   # env override: A/B builds during development

# enums (include/gasal_b200.h, reference src/gasal.h:37-72)
WITHOUT_START, WITH_START, WITH_TB = 0, 1, 2
FALSE, TRUE = 0, 1
NONE, QUERY, TARGET, BOTH = 0, 1, 2, 3
UNKNOWN, GLOBAL, SEMI_GLOBAL, LOCAL, MICROLOCAL, BANDED, KSW = 0, 1, 2, 3, 4, 5, 6


class HostBatch(C.Structure):
    pass


HostBatch._fields_ = [("data", C.POINTER(C.c_uint8)), ("page_size", C.c_uint32), ("data_size", C.c_uint32), ("offset", C.c_uint32),
                      ("is_locked", C.c_int), ("next", C.POINTER(HostBatch))]


class GasalRes(C.Structure):
    _fields_ = [("aln_score", C.POINTER(C.c_int32)), ("query_batch_end", C.POINTER(C.c_int32)), ("target_batch_end", C.POINTER(C.c_int32)),
                ("query_batch_start", C.POINTER(C.c_int32)), ("target_batch_start", C.POINTER(C.c_int32)),
                ("cigar", C.POINTER(C.c_uint8)), ("n_cigar_ops", C.POINTER(C.c_uint32))]


class GpuStorage(C.Structure):
    _fields_ = [
        ("unpacked_query_batch", C.c_void_p), ("unpacked_target_batch", C.c_void_p),
        ("packed_query_batch", C.c_void_p), ("packed_target_batch", C.c_void_p),
        ("query_batch_offsets", C.c_void_p), ("target_batch_offsets", C.c_void_p),
        ("query_batch_lens", C.c_void_p), ("target_batch_lens", C.c_void_p),
        ("host_seed_scores", C.POINTER(C.c_uint32)), ("seed_scores", C.c_void_p),
        ("extensible_host_unpacked_query_batch", C.POINTER(HostBatch)), ("extensible_host_unpacked_target_batch", C.POINTER(HostBatch)),
        ("host_query_op", C.POINTER(C.c_uint8)), ("host_target_op", C.POINTER(C.c_uint8)),
        ("query_op", C.c_void_p), ("target_op", C.c_void_p),
        ("host_query_batch_offsets", C.POINTER(C.c_uint32)), ("host_target_batch_offsets", C.POINTER(C.c_uint32)),
        ("host_query_batch_lens", C.POINTER(C.c_uint32)), ("host_target_batch_lens", C.POINTER(C.c_uint32)),
        ("host_res", C.POINTER(GasalRes)), ("device_cpy", C.POINTER(GasalRes)), ("device_res", C.c_void_p),
        ("host_res_second", C.POINTER(GasalRes)), ("device_res_second", C.c_void_p), ("device_cpy_second", C.POINTER(GasalRes)),
        ("gpu_max_query_batch_bytes", C.c_uint32), ("gpu_max_target_batch_bytes", C.c_uint32),
        ("host_max_query_batch_bytes", C.c_uint32), ("host_max_target_batch_bytes", C.c_uint32),
        ("gpu_max_n_alns", C.c_uint32), ("host_max_n_alns", C.c_uint32), ("current_n_alns", C.c_uint32),
        ("packed_tb_matrix_size", C.c_uint64), ("packed_tb_matrices", C.c_void_p),
        ("str", C.c_void_p), ("is_free", C.c_int), ("id", C.c_int),
        ("b200_priv", C.c_void_p),
    ]


class GpuStorageV(C.Structure):
    _fields_ = [("n", C.c_int), ("a", C.POINTER(GpuStorage))]


class SubstScores(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32)]


class Sched(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("gpu_ids", C.POINTER(C.c_int32)), ("streams_per_gpu", C.c_int32), ("max_batch_pairs", C.c_uint32),
                ("max_batch_bytes", C.c_uint64), ("max_batch_cells", C.c_uint64), ("keep_input_order", C.c_int32), ("host_threads_per_gpu", C.c_int32)]


class Results(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_int32)) for n in ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start", "target_batch_start",
                                                    "aln_score2", "query_batch_end2", "target_batch_end2")] + \
               [("cigar", C.POINTER(C.c_uint8)), ("n_cigar_ops", C.POINTER(C.c_uint32))]


class Params(C.Structure):
    _fields_ = [("algo", C.c_int32), ("start_pos", C.c_int32), ("secondBest", C.c_int32), ("isPacked", C.c_int32),
                ("isReverseComplement", C.c_int32), ("semiglobal_skipping_head", C.c_int32), ("semiglobal_skipping_tail", C.c_int32),
                ("k_band", C.c_int32)]


# every symbol include/gasal_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "gasal_params_default", "gasal_init_gpu_storage_v", "gasal_init_streams", "gasal_destroy_streams", "gasal_destroy_gpu_storage_v",
    "gasal_gpu_mem_alloc", "gasal_gpu_mem_free",
    "gasal_copy_subst_scores", "gasal_aln_async", "gasal_is_aln_async_done",
    "gasal_host_batch_new", "gasal_host_batch_destroy", "gasal_host_batch_getlast", "gasal_host_batch_reset", "gasal_host_batch_fill",
    "gasal_host_batch_add", "gasal_host_batch_addbase", "gasal_host_batch_print", "gasal_host_batch_printall",
    "gasal_host_alns_resize", "gasal_op_fill", "gasal_set_device",
    "gasal_res_new_host", "gasal_res_new_device", "gasal_res_new_device_cpy", "gasal_res_destroy_host", "gasal_res_destroy_device",
    "gasalx_set_n_code", "gasalx_set_kernel_policy", "gasalx_set_revcomp_mode", "gasalx_set_coordinate_fix", "gasalx_set_host_packing", "gasalx_storage_host_packed",
    "gasalx_set_host_write_combined", "gasalx_set_fast_path_passes", "gasalx_span_ms", "gasalx_last_batch_ms", "gasalx_kernel_launches", "gasalx_last_kernel_name",
    "gasalx_aln_device_resident", "gasalx_host_batch_rearm", "gasalx_version",
    "gasalx_sched_default", "gasalx_plan_batches", "gasalx_align_all", "gasalx_batcher_release",
    "gasalx_read_fasta_pairs", "gasalx_free_fasta_pairs", "gasalx_pack_host", "gasalx_pack_host_at", "gasalx_cigar_string",
    "gasalx_cigar_text_async", "gasalx_sam_record",
    "gasalx_try_aln_async", "gasalx_try_init_streams", "gasalx_last_error",
]


class FastaPairs(C.Structure):
    _fields_ = [("n", C.c_uint64), ("q", C.POINTER(C.c_uint8)), ("t", C.POINTER(C.c_uint8)), ("q_bytes", C.c_uint64), ("t_bytes", C.c_uint64),
                ("q_offs", C.POINTER(C.c_uint64)), ("t_offs", C.POINTER(C.c_uint64)), ("q_lens", C.POINTER(C.c_uint32)), ("t_lens", C.POINTER(C.c_uint32)),
                ("q_op", C.POINTER(C.c_uint8)), ("t_op", C.POINTER(C.c_uint8)), ("q_names", C.c_void_p), ("t_names", C.c_void_p),
                ("max_query_len", C.c_uint32), ("max_target_len", C.c_uint32), ("any_op", C.c_int)]


_lib = None


def load() -> C.CDLL:
    """Load libgasal_b200.so (no compute happens at load time). Raises if the native library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"native engine missing: {LIB_PATH} -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.gasal_params_default.argtypes = [P(Params)]
    lib.gasal_init_gpu_storage_v.argtypes = [C.c_int]; lib.gasal_init_gpu_storage_v.restype = GpuStorageV
    lib.gasal_init_streams.argtypes = [P(GpuStorageV), C.c_int, C.c_int, C.c_int, P(Params)]
    lib.gasal_destroy_streams.argtypes = [P(GpuStorageV), P(Params)]
    lib.gasal_destroy_gpu_storage_v.argtypes = [P(GpuStorageV)]
    lib.gasal_copy_subst_scores.argtypes = [P(SubstScores)]
    lib.gasal_aln_async.argtypes = [P(GpuStorage), C.c_uint32, C.c_uint32, C.c_uint32, P(Params)]
    lib.gasal_is_aln_async_done.argtypes = [P(GpuStorage)]; lib.gasal_is_aln_async_done.restype = C.c_int
    lib.gasal_host_batch_fill.argtypes = [P(GpuStorage), C.c_uint32, C.c_char_p, C.c_uint32, C.c_int]; lib.gasal_host_batch_fill.restype = C.c_uint32
    lib.gasal_host_batch_add.argtypes = [P(GpuStorage), C.c_uint32, C.c_char_p, C.c_uint32, C.c_int]; lib.gasal_host_batch_add.restype = C.c_uint32
    lib.gasal_host_batch_reset.argtypes = [P(GpuStorage)]
    lib.gasal_host_alns_resize.argtypes = [P(GpuStorage), C.c_int, P(Params)]
    lib.gasal_op_fill.argtypes = [P(GpuStorage), P(C.c_uint8), C.c_uint32, C.c_int]
    lib.gasal_set_device.argtypes = [C.c_int, C.c_int]
    lib.gasalx_set_n_code.argtypes = [C.c_int, C.c_int]
    lib.gasalx_set_kernel_policy.argtypes = [C.c_int]
    lib.gasalx_set_coordinate_fix.argtypes = [C.c_int]
    lib.gasalx_set_revcomp_mode.argtypes = [C.c_int]
    lib.gasalx_set_host_packing.argtypes = [C.c_int]
    lib.gasalx_set_host_write_combined.argtypes = [C.c_int]
    lib.gasalx_set_fast_path_passes.argtypes = [C.c_int]
    lib.gasalx_storage_host_packed.argtypes = [P(GpuStorage)]; lib.gasalx_storage_host_packed.restype = C.c_int
    lib.gasalx_span_ms.argtypes = [P(P(GpuStorage)), C.c_int, P(C.c_float)]; lib.gasalx_span_ms.restype = C.c_int
    lib.gasal_gpu_mem_alloc.argtypes = [P(GpuStorage), C.c_int, C.c_int, C.c_int, P(Params)]
    lib.gasal_gpu_mem_free.argtypes = [P(GpuStorage), P(Params)]
    lib.gasal_res_new_device_cpy.argtypes = [C.c_uint32, P(Params)]; lib.gasal_res_new_device_cpy.restype = P(GasalRes)
    lib.gasal_host_batch_addbase.argtypes = [P(GpuStorage), C.c_uint32, C.c_char, C.c_int]; lib.gasal_host_batch_addbase.restype = C.c_uint32
    lib.gasalx_last_batch_ms.argtypes = [P(GpuStorage), P(C.c_float)]; lib.gasalx_last_batch_ms.restype = C.c_int
    lib.gasalx_kernel_launches.argtypes = [P(GpuStorage)]; lib.gasalx_kernel_launches.restype = C.c_uint64
    lib.gasalx_last_kernel_name.argtypes = [P(GpuStorage)]; lib.gasalx_last_kernel_name.restype = C.c_char_p
    lib.gasalx_aln_device_resident.argtypes = [P(GpuStorage), C.c_uint32, C.c_uint32, C.c_uint32, P(Params)]
    lib.gasalx_host_batch_rearm.argtypes = [P(GpuStorage), C.c_uint32, C.c_uint32, C.c_uint32]
    lib.gasalx_version.restype = C.c_char_p
    lib.gasalx_sched_default.argtypes = [P(Sched)]
    lib.gasalx_plan_batches.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, P(Sched), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.gasalx_plan_batches.restype = C.c_int64
    lib.gasalx_align_all.argtypes = [C.c_void_p] * 9 + [C.c_uint64, P(Params), P(Sched), P(Results), C.c_void_p]
    lib.gasalx_align_all.restype = C.c_int
    lib.gasalx_read_fasta_pairs.argtypes = [C.c_char_p, C.c_char_p, C.c_int, P(FastaPairs)]; lib.gasalx_read_fasta_pairs.restype = C.c_int
    lib.gasalx_free_fasta_pairs.argtypes = [P(FastaPairs)]
    lib.gasalx_pack_host.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    lib.gasalx_pack_host_at.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.gasalx_cigar_string.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32]; lib.gasalx_cigar_string.restype = C.c_int
    lib.gasalx_cigar_text_async.argtypes = [P(GpuStorage), C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]; lib.gasalx_cigar_text_async.restype = C.c_int
    lib.gasalx_sam_record.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_int, C.c_char_p, C.c_int64, C.c_int, C.c_char_p, C.c_char_p, C.c_uint32,
                                      C.c_char_p, C.c_int32, C.c_int32]
    lib.gasalx_sam_record.restype = C.c_int
    lib.gasalx_try_aln_async.argtypes = [P(GpuStorage), C.c_uint32, C.c_uint32, C.c_uint32, P(Params)]; lib.gasalx_try_aln_async.restype = C.c_int
    lib.gasalx_try_init_streams.argtypes = [P(GpuStorageV), C.c_int, C.c_int, C.c_int, P(Params)]; lib.gasalx_try_init_streams.restype = C.c_int
    lib.gasalx_last_error.restype = C.c_char_p
    _lib = lib
    return lib
