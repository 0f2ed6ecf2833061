"""ctypes binding of libgrafimo_b200.so -- the C ABI declared in include/grafimo_b200.h.

This is the binding a GRAFIMO maintainer would add.  There is no CPU fallback:
if the shared library is missing, or no CUDA device is present when a compute entry point is
called, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrafimo_b200.so")

GB2_OK = 0
GB2_ERR_CAPACITY = 4
NARROW_WIDTH = 32  # widest k-mer that fits one packed word
MAX_WIDTH = 64
RANGE = 1000
NCCL_ID_BYTES = 128


class GrafimoB200Error(RuntimeError):
    """An entry point of the CUDA library returned a non-zero status."""

    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: {detail or 'error'} (gb2 status {code})")


class Hit(ctypes.Structure):
    _fields_ = [("row", ctypes.c_uint64), ("score", ctypes.c_int32), ("strand", ctypes.c_uint32)]


class Report(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_uint64), ("index_base", ctypes.c_uint64), ("width", ctypes.c_int32), ("layout", ctypes.c_int32),
                ("want_q", ctypes.c_int32), ("reserved", ctypes.c_int32), ("d_kmer", ctypes.c_void_p), ("d_strand", ctypes.c_void_p),
                ("d_start", ctypes.c_void_p), ("d_stop", ctypes.c_void_p), ("d_freq", ctypes.c_void_p), ("d_ref", ctypes.c_void_p),
                ("d_bin", ctypes.c_void_p), ("d_name", ctypes.c_void_p), ("d_strings", ctypes.c_void_p),
                ("d_string_off", ctypes.c_void_p), ("first_score", ctypes.c_int32), ("first_p", ctypes.c_int32),
                ("first_q", ctypes.c_int32), ("first_name", ctypes.c_int32), ("first_chrom", ctypes.c_int32),
                ("first_const", ctypes.c_int32)]


class GraphInput(ctypes.Structure):
    """gb2_graph_input (include/grafimo_b200.h)."""
    _fields_ = [("h_ref", ctypes.c_void_p), ("ref_len", ctypes.c_int64), ("n_variants", ctypes.c_int64),
                ("h_var_pos", ctypes.c_void_p), ("h_var_ref_len", ctypes.c_void_p), ("h_alt_off", ctypes.c_void_p),
                ("h_alt", ctypes.c_void_p), ("n_hap", ctypes.c_int32), ("words", ctypes.c_int32),
                ("h_gt_bits", ctypes.c_void_p), ("max_node_len", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class GraphInfo(ctypes.Structure):
    _fields_ = [("n_nodes", ctypes.c_int64), ("n_edges", ctypes.c_int64), ("n_bases", ctypes.c_int64),
                ("n_sets", ctypes.c_int64), ("n_hap", ctypes.c_int32), ("words", ctypes.c_int32)]


class MotifInfo(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("n_chunks", ctypes.c_int32), ("lut_replicas", ctypes.c_int32),
        ("monotone", ctypes.c_int32), ("lo", ctypes.c_int64), ("hi", ctypes.c_int64), ("span", ctypes.c_int64),
        ("min_val", ctypes.c_int64), ("scale", ctypes.c_int64), ("offset", ctypes.c_double),
        ("total", ctypes.c_double), ("smem_bytes", ctypes.c_int64), ("chunk_bases", ctypes.c_int32),
        ("hist_global", ctypes.c_int32),
    ]


_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_int = ctypes.c_int
_dbl = ctypes.c_double

# name -> (restype, argtypes); every symbol include/grafimo_b200.h declares
SIGNATURES = {
    "gb2_abi_version": (_int, []),
    "gb2_error_string": (ctypes.c_char_p, [_int]),
    "gb2_ctx_create": (_int, [_int, _vp, ctypes.POINTER(_vp)]),
    "gb2_ctx_destroy": (_int, [_vp]),
    "gb2_ctx_set_stream": (_int, [_vp, _vp]),
    "gb2_ctx_sync": (_int, [_vp]),
    "gb2_ctx_last_error": (ctypes.c_char_p, [_vp]),
    "gb2_ctx_launch_count": (_i64, [_vp]),
    "gb2_device_count": (_int, []),
    "gb2_ctx_sm_count": (_int, [_vp]),
    "gb2_encode_kmers": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _vp, _vp]),
    "gb2_tsv_index_lines": (_int, [_vp, _vp, _i64, _int, _vp, _u64, _vp]),
    "gb2_tsv_parse_rows": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_pval_dp_batched": (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    "gb2_motif_create": (_int, [_vp, _vp, _int, _vp, _i64, _i64, _dbl, ctypes.POINTER(_vp)]),
    "gb2_motif_create_batched": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_motif_destroy": (_int, [_vp]),
    "gb2_motif_get_info": (_int, [_vp, ctypes.POINTER(MotifInfo)]),
    "gb2_motif_get_ptable": (_int, [_vp, _vp, _vp]),
    "gb2_motif_ptable_device": (_vp, [_vp]),
    "gb2_score": (_int, [_vp, _vp, _vp, _vp, _i64, _u64, _int, _dbl, _vp, _vp, _u64, _vp, _vp]),
    "gb2_qvalues_from_hist": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_qvalues_from_hist_many": (_int, [_vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_bh_pvalues": (_int, [_vp, _vp, _i64, _vp]),
    "gb2_finalize_hits": (_int, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _dbl, _int, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_finalize_hits_many": (_int, [_vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _u64, _u64, _dbl, _int, _dbl, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp]),
    "gb2_finalize_dense": (_int, [_vp, _vp, _vp, _u64, _int, _u64, _vp, _vp, _dbl, _int, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_tally_haplotypes": (_int, [_vp, _vp, _vp, _i64, _vp, _u64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "gb2_graph_create": (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32,
                                _i64, _vp, ctypes.POINTER(_vp)]),
    "gb2_graph_destroy": (_int, [_vp]),
    "gb2_graph_build": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp, ctypes.c_int32,
                               ctypes.POINTER(_vp)]),
    "gb2_graph_build_stats": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp, ctypes.c_int32,
                                     ctypes.c_int32, _i64, _vp]),
    "gb2_graph_build_batch": (_int, [_vp, ctypes.c_int32, _vp, ctypes.c_int32, ctypes.POINTER(_vp)]),
    "gb2_graph_get_info": (_int, [_vp, _vp]),
    "gb2_graph_prepare": (_int, [_vp, _vp, ctypes.c_int32, _vp, _vp, _int, ctypes.POINTER(_u64)]),
    "gb2_graph_extract": (_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_vcf_parse_fields": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_vcf_parse_genotypes": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _int, ctypes.c_int32, ctypes.c_int32, _vp, _vp]),
    "gb2_report_measure": (_int, [_vp, _vp, _vp, ctypes.POINTER(_u64)]),
    "gb2_report_write": (_int, [_vp, _vp, _vp, _vp, _u64]),
    "gb2_scan_host": (_int, [_vp, _vp, _vp, _i64, _int, _i64, _int, _dbl, _int, _int, _u64, _vp, _vp, _vp, _vp, _vp,
                             _vp, ctypes.POINTER(_u64), _vp]),
    "gb2_scan_host_packed": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _dbl, _int, _int, _u64, _vp, _vp, _vp, _vp, _vp, _vp,
                                    ctypes.POINTER(_u64), _vp]),
    "gb2_comm_unique_id": (_int, [_vp]),
    "gb2_comm_init": (_int, [_vp, _vp, _int, _int]),
    "gb2_comm_destroy": (_int, [_vp]),
    "gb2_comm_info": (_int, [_vp, ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "gb2_allreduce_hist": (_int, [_vp, _vp, _i64]),
    "gb2_allreduce_max_f64": (_int, [_vp, _vp, _i64]),
    "gb2_allgather_bytes": (_int, [_vp, _vp, _vp, _i64]),
    "gb2_host_alloc": (_int, [_u64, ctypes.POINTER(_vp)]),
    "gb2_host_free": (_int, [_vp]),
    "gb2_encode_sequences": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gb2_score_sequences": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _u64, _int, _dbl, _vp, _vp, _u64, _vp, _vp,
                                   ctypes.POINTER(_u64)]),
    "gb2_pack_sequence_host": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "gb2_scan_last_transfer": (_int, [_vp, ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "gb2_scan_host_sequences": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _vp, _vp, _int, _dbl, _int, _int, _u64, _vp, _vp, _vp,
                                       _vp, _vp, _vp, ctypes.POINTER(_u64), _vp]),
}

_lib = None


def load():
    """Loads the CUDA library (once).  Raises if it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GrafimoB200Error(
            -1, "grafimo_b200", f"{LIB_PATH} is missing: build it with `python -m grafimo_b200.build` "
            "(needs nvcc); this package has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.gb2_abi_version() != 3:
        raise GrafimoB200Error(-1, "grafimo_b200", "ABI version mismatch between _lib.py and the shared library")
    _lib = lib
    return lib


def check(code, where, ctx=None):
    if code == GB2_OK:
        return
    lib = load()
    detail = ""
    if ctx:
        detail = lib.gb2_ctx_last_error(ctx).decode("utf-8", "replace")
    if not detail:
        detail = lib.gb2_error_string(code).decode()
    raise GrafimoB200Error(code, where, detail)
