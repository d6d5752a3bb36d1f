"""ctypes binding of libpgstrom_cuda.so (include/pgstrom_cuda.h).

The library is the product; this module only declares prototypes.  It fails
loudly when the shared object is missing: there is no Python or CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, "libpgstrom_cuda.so")


class kern_colmeta(C.Structure):
    _fields_ = [("attbyval", C.c_int8), ("attalign", C.c_int8),
                ("attlen", C.c_int16), ("attnum", C.c_int16),
                ("attcacheoff", C.c_int16)]


class kern_data_store(C.Structure):
    _fields_ = [("hostptr", C.c_uint64), ("length", C.c_uint32),
                ("usage", C.c_uint32), ("ncols", C.c_uint32),
                ("nitems", C.c_uint32), ("nrooms", C.c_uint32),
                ("nblocks", C.c_uint32), ("maxblocks", C.c_uint32),
                ("format", C.c_int8), ("tdhasoid", C.c_int8),
                ("tdtypeid", C.c_uint32), ("tdtypmod", C.c_int32)]


class pgs_session_config(C.Structure):
    _fields_ = [("device", C.c_int), ("needs_grouping", C.c_int),
                ("num_groups", C.c_double), ("max_async_chunks", C.c_int),
                ("max_chunk_rows", C.c_uint32), ("max_chunk_bytes", C.c_size_t),
                ("result_ncols", C.c_int),
                ("result_colmeta", C.POINTER(kern_colmeta))]


class pgs_bulkslot(C.Structure):
    pass


RELEASE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
pgs_bulkslot._fields_ = [("kds", C.c_void_p), ("krowmap", C.c_void_p),
                         ("release", RELEASE_FN), ("release_arg", C.c_void_p)]
BULK_EXEC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(pgs_bulkslot))

# every symbol include/pgstrom_cuda.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "pgstrom_strerror": (C.c_char_p, [C.c_int]),
    "pgs_last_error": (C.c_char_p, []),
    "pgstrom_abi_version": (C.c_int, []),
    "pgstrom_guc_set": (C.c_int, [C.c_char_p, C.c_char_p]),
    "pgstrom_guc_get": (C.c_char_p, [C.c_char_p]),
    "pgstrom_guc_list_json": (C.c_char_p, []),
    "pgstrom_guc_reset_all": (None, []),
    "pgstrom_grafter_json": (C.c_void_p, [C.c_char_p]),
    "pgs_plan_free": (None, [C.c_void_p]),
    "pgs_plan_tree_json": (C.c_char_p, [C.c_void_p]),
    "pgs_plan_explain": (C.c_char_p, [C.c_void_p, C.c_int]),
    "pgs_plan_num_gpupreagg": (C.c_int, [C.c_void_p]),
    "pgs_plan_reject_reason": (C.c_char_p, [C.c_void_p]),
    "pgs_plan_kernel_source": (C.c_char_p, [C.c_void_p, C.c_int]),
    "pgs_plan_extra_flags": (C.c_int, [C.c_void_p, C.c_int]),
    "pgs_plan_kparams": (C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]),
    "pgs_plan_needs_grouping": (C.c_int, [C.c_void_p, C.c_int]),
    "pgs_plan_num_groups": (C.c_double, [C.c_void_p, C.c_int]),
    "pgs_plan_describe_json": (C.c_char_p, [C.c_void_p, C.c_int]),
    "pgs_plan_result_colmeta": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(kern_colmeta), C.c_int]),
    "pgstrom_codegen_available_expression_json": (C.c_int, [C.c_char_p]),
    "pgstrom_kds_head_length": (C.c_size_t, [C.c_int]),
    "pgstrom_kds_column_length": (C.c_size_t, [C.c_int, C.POINTER(kern_colmeta), C.c_uint32,
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "pgstrom_kds_column_build": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int,
                                           C.POINTER(kern_colmeta), C.c_uint32,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "pgstrom_kds_row_length": (C.c_size_t, [C.c_int, C.c_uint32, C.c_uint32]),
    "pgstrom_kds_row_init": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int,
                                       C.POINTER(kern_colmeta), C.c_uint32, C.c_uint32]),
    "pgstrom_kds_row_insert_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "pgstrom_kds_flat_init": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int,
                                        C.POINTER(kern_colmeta), C.c_uint32]),
    "pgstrom_kds_flat_insert_tuple": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "pgstrom_colmeta_set_cacheoff": (None, [C.c_int, C.POINTER(kern_colmeta)]),
    "pgstrom_heap_form_pages": (C.c_long, [C.c_int, C.POINTER(kern_colmeta), C.c_uint32,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgstrom_kds_tupslot_length": (C.c_size_t, [C.c_int, C.c_uint32]),
    "pgstrom_kds_tupslot_init": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int,
                                           C.POINTER(kern_colmeta), C.c_uint32]),
    "pgstrom_fetch_data_store": (C.c_int, [C.c_void_p, C.c_uint32,
                                           C.POINTER(C.c_uint64), C.c_char_p]),
    "pgstrom_fixup_kernel_numeric": (C.c_int, [C.c_uint64, C.c_char_p, C.c_size_t]),
    "pgstrom_fixup_kernel_text": (C.c_size_t, [C.c_uint64, C.c_int, C.c_void_p, C.c_size_t]),
    "pgstrom_fixup_kernel_text_heap": (C.c_size_t, [C.c_uint64, C.c_int, C.c_void_p, C.c_size_t,
                                                    C.c_void_p, C.c_size_t]),
    "pgstrom_numeric_from_text": (C.c_size_t, [C.c_char_p, C.c_void_p, C.c_size_t]),
    "pgstrom_numeric_to_text": (C.c_size_t, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "pgs_program_build": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_char_p)]),
    "pgs_program_release": (None, [C.c_void_p]),
    "pgs_program_cubin": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "pgs_program_info_json": (C.c_char_p, []),
    "pgs_cuda_init": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "pgs_cuda_device_count": (C.c_int, []),
    "pgs_cuda_device_info_json": (C.c_char_p, []),
    "pgs_cuda_shutdown": (None, []),
    "pgs_chunk_alloc": (C.c_void_p, [C.c_size_t]),
    "pgs_chunk_free": (None, [C.c_void_p]),
    "pgs_preagg_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(pgs_session_config),
                                  C.POINTER(C.c_void_p)]),
    "pgs_preagg_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "pgs_preagg_submit_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                           C.c_void_p, C.POINTER(C.c_int64)]),
    "pgs_preagg_submit_device_format": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                                  C.c_int, C.c_void_p, C.POINTER(C.c_int64)]),
    "pgs_preagg_wait": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int32)]),
    "pgs_preagg_recheck_rows": (C.c_int64, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint32), C.c_int64]),
    "pgs_preagg_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_int32)]),
    "pgs_nccl_get_unique_id": (C.c_int, [C.c_void_p]),
    "pgs_nccl_comm_init_rank": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int,
                                          C.POINTER(C.c_void_p)]),
    "pgs_nccl_comm_destroy": (None, [C.c_void_p]),
    "pgs_preagg_merge_nccl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pgs_preagg_peer_setup": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pgs_preagg_peer_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pgs_preagg_peer_attach_session": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pgs_preagg_merge_peer": (C.c_int, [C.c_void_p]),
    "pgs_preagg_merge_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pgs_preagg_state_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)]),
    "pgs_preagg_state_import": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "pgs_preagg_state_reset": (C.c_int, [C.c_void_p]),
    "pgs_preagg_perfmon_json": (C.c_char_p, [C.c_void_p]),
    "pgs_preagg_abort": (None, [C.c_void_p]),
    "pgs_preagg_close": (None, [C.c_void_p]),
    "pgs_preagg_stream": (C.c_void_p, [C.c_void_p]),
    "pgs_preagg_key_heap": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "pgs_preagg_launch_count": (C.c_uint64, [C.c_void_p]),
    "pgs_device_alloc": (C.c_void_p, [C.c_int, C.c_size_t]),
    "pgs_device_free": (None, [C.c_int, C.c_void_p]),
    "pgs_device_upload": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pgs_device_l2_flush": (C.c_int, [C.c_int]),
    "gpupreagg_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, BULK_EXEC_FN, C.c_void_p,
                                  C.POINTER(C.c_void_p)]),
    "gpupreagg_exec": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_char_p]),
    "gpupreagg_key_heap": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "gpupreagg_recheck_rows": (C.c_int64, [C.c_void_p, C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_uint32), C.c_int64]),
    "gpupreagg_recheck_chunk": (C.c_void_p, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "gpupreagg_recheck_done": (C.c_int, [C.c_void_p, C.c_uint32]),
    "gpupreagg_end": (C.c_char_p, [C.c_void_p]),
    "gpupreagg_rescan": (C.c_int, [C.c_void_p]),
    "gpupreagg_explain": (C.c_char_p, [C.c_void_p, C.c_int, C.c_int]),
    # 7. SQL-side functions (host only)
    "pgs_partial_nrows": (C.c_int32, [C.c_int, C.c_char_p, C.c_char_p]),
    "pgs_psum_x2_float8": (C.c_int, [C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "pgs_pcov_float8": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                  C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "pgs_avg_int8_accum": (C.c_int, [C.POINTER(C.c_int64), C.c_int32, C.c_int64]),
    "pgs_sum_int8_accum": (C.c_int, [C.POINTER(C.c_int64), C.c_int64]),
    "pgs_sum_int8_final": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pgs_sum_float8_accum": (C.c_int, [C.POINTER(C.c_double), C.c_int32, C.c_double]),
    "pgs_variance_float8_accum": (C.c_int, [C.POINTER(C.c_double), C.c_int32,
                                            C.c_double, C.c_double]),
    "pgs_covariance_float8_accum": (C.c_int, [C.POINTER(C.c_double), C.c_int32,
                                              C.POINTER(C.c_double)]),
    "pgs_numeric_avg_init": (C.c_void_p, []),
    "pgs_numeric_avg_free": (None, [C.c_void_p]),
    "pgs_numeric_avg_accum": (C.c_int, [C.c_void_p, C.c_int32, C.c_int, C.c_char_p]),
    "pgs_numeric_avg_count": (C.c_int64, [C.c_void_p]),
    "pgs_numeric_avg_sum_text": (C.c_size_t, [C.c_void_p, C.c_char_p, C.c_size_t]),
}

_lib = None


def load():
    """Loads the shared library (building is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise RuntimeError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no fallback implementation)" % LIBPATH)
    lib = C.CDLL(LIBPATH, mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class StromError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = load().pgstrom_strerror(code).decode()
        if detail:
            msg += ": " + detail
        super().__init__("[StromError %d] %s" % (code, msg))


def check(rc):
    if rc != 0:
        raise StromError(rc, load().pgs_last_error().decode(errors="replace"))
    return rc
