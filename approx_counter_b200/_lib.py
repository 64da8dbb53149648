"""ctypes binding of libapc.so (include/apc.h).

The shared library is built in-tree (approx_counter_b200/csrc/libapc.so) by
`__graft_entry__.build()` / `make -C approx_counter_b200/csrc`.  There is no
fallback: if the library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# APC_LIB_PATH: development override used by tools/ to A/B differently compiled builds
LIB_PATH = os.environ.get("APC_LIB_PATH") or os.path.join(_HERE, "csrc", "libapc.so")

APC_OK = 0
ERRORS = {
    -1: "APC_ERR_INVALID", -2: "APC_ERR_CUDA", -3: "APC_ERR_NO_DEVICE", -4: "APC_ERR_NO_SAMPLE",
    -5: "APC_ERR_NO_QUERIES", -6: "APC_ERR_NOMEM", -7: "APC_ERR_CAPACITY", -8: "APC_ERR_COMM",
    -9: "APC_ERR_FORMAT",
}

# every symbol include/apc.h declares: (restype, argtypes)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p


class ApcTiming(C.Structure):
    _fields_ = [("upload_ms", C.c_float), ("exact_ms", C.c_float), ("scan_ms", C.c_float),
                ("total_ms", C.c_float), ("scan_launches", C.c_uint64),
                ("exact_launches", C.c_uint64)]


class ApcScanStats(C.Structure):
    _fields_ = [("scans", C.c_uint64), ("lop3_executed", C.c_double), ("lop3_top", C.c_double),
                ("lop3_planned", C.c_double), ("lop3_one_kmer_per_warp", C.c_double)]


SYMBOLS = {
    "apc_version": (C.c_int, []),
    "apc_strerror": (C.c_char_p, [C.c_int]),
    "apc_device_count": (C.c_int, []),
    "apc_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "apc_destroy": (None, [_vp]),
    "apc_last_error": (C.c_char_p, [_vp]),
    "apc_set_stream": (C.c_int, [_vp, _vp]),
    "apc_sync": (C.c_int, [_vp]),
    "apc_reserve": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_uint8, C.c_uint32]),
    "apc_upload_sample": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32]),
    "apc_upload_sample_async": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32]),
    "apc_upload_sample_ragged": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "apc_sample_info": (C.c_int, [_vp, _u64p, C.POINTER(C.c_uint32), _u64p]),
    "apc_ingest_fastx": (C.c_int, [_vp, _vp, C.c_uint64, _u64p, C.POINTER(C.c_int)]),
    "apc_ingest_lengths": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "apc_sample_resident": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _u64p]),
    "apc_upload_sample_peer": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64]),
    "apc_download_sample": (C.c_int, [_vp, _vp, C.c_uint64]),
    "apc_ingest_timing": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "apc_exact_topn": (C.c_int, [_vp, C.c_uint8, C.c_float, C.c_uint64, _vp, C.c_uint64,
                                 _vp, _vp, _u64p, _u64p, _u64p]),
    "apc_exact_solid": (C.c_int, [_vp, C.c_uint8, C.c_float, C.c_uint64, _vp, C.c_uint64,
                                  _vp, _vp, C.c_uint64, _u64p, _u64p, _u64p]),
    "apc_approx_count": (C.c_int, [_vp, C.c_uint8, _vp, C.c_uint32, _vp]),
    "apc_approx_count_async": (C.c_int, [_vp, C.c_uint8, _vp, C.c_uint32, _vp]),
    "apc_set_queries": (C.c_int, [_vp, C.c_uint8, _vp, C.c_uint32]),
    "apc_plan_queries": (C.c_int, [C.c_uint8, _vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "apc_scan": (C.c_int, [_vp, _vp]),
    "apc_get_counts": (C.c_int, [_vp, _vp]),
    "apc_counts_device_ptr": (_vp, [_vp]),
    "apc_last_scan_launches": (C.c_uint64, [_vp]),
    "apc_last_timing": (C.c_int, [_vp, C.POINTER(ApcTiming)]),
    "apc_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "apc_measure_int_peak": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]),
    "apc_microbench": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_double)]),
    "apc_comm_unique_id": (C.c_int, [_vp]),
    "apc_comm_init_rank": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "apc_comm_destroy": (C.c_int, [_vp]),
    "apc_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "apc_allreduce_counts": (C.c_int, [_vp, _vp, C.c_uint64]),
    "apc_scan_allreduce": (C.c_int, [_vp, _vp]),
    "apc_scan_stats_read": (C.c_int, [_vp, C.POINTER(ApcScanStats)]),
}

_lib = None


class ApcError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        name = ERRORS.get(status, str(status))
        super().__init__(f"{name}: {detail}" if detail else name)


def load():
    """dlopen libapc.so and type every exported symbol.  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C approx_counter_b200/csrc` (no CPU fallback exists)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
