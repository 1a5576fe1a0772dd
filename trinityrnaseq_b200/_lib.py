"""ctypes binding of libtrinity_gpu.so (the C ABI declared in include/trinity_gpu.h).

The library is the product; there is no Python or CPU implementation behind these calls.  If the shared
object is missing the import fails loudly, and if no B200 is visible tg_init fails with TG_ERR_NOGPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtrinity_gpu.so")

TG_OK, TG_ERR_CUDA, TG_ERR_ARG, TG_ERR_NOMEM, TG_ERR_TABLE, TG_ERR_NOGPU = 0, -1, -2, -3, -4, -5
TG_TABLE_COUNT, TG_TABLE_LABEL = 0, 1
TG_HISTO_BINS = 10002
TG_IPC_HANDLE_BYTES = 64

# every symbol include/trinity_gpu.h declares: name -> (restype, argtypes)
_u64, _u32, _i32, _vp, _cp = C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_char_p
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "tg_version": (_i32, []),
    "tg_last_error": (_cp, []),
    "tg_device_count": (_i32, []),
    "tg_init": (_i32, [_i32, _pp]),
    "tg_destroy": (None, [_vp]),
    "tg_device_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_u64), C.POINTER(_u64)]),
    "tg_sync": (_i32, [_vp]),
    "tg_log_overflow_check": (_i32, [_vp, C.POINTER(_i32)]),
    "tg_launch_count": (_u64, [_vp]),
    "tg_ctx_set": (_i32, [_vp, _cp, _cp]),
    "tg_kernel_times": (_i32, [_vp, _vp, _u64]),
    "tg_host_alloc": (_vp, [_u64]),
    "tg_host_free": (None, [_vp]),
    "tg_free": (None, [_vp]),
    "tg_table_create": (_i32, [_vp, _i32, _i32, _u64, _pp]),
    "tg_table_destroy": (None, [_vp]),
    "tg_table_reserve": (_i32, [_vp, _u64]),
    "tg_table_info": (_i32, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "tg_table_clear": (_i32, [_vp]),
    "tg_table_create_sharded": (_i32, [_vp, _i32, _i32, _u64, _u32, _u32, _u32, _pp]),
    "tg_table_geometry": (_i32, [_vp, C.POINTER(_u64), C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "tg_table_resize": (_i32, [_vp, _u64]),
    "tg_table_count_min": (_i32, [_vp, _u32, C.POINTER(_u64)]),
    "tg_table_compact_into": (_i32, [_vp, _u32, _vp]),
    "tg_table_slots_dev": (_i32, [_vp, _pp, C.POINTER(_u64)]),
    "tg_table_set_distinct": (_i32, [_vp, _u64]),
    "tg_count_reads": (_i32, [_vp, _vp, _u64, _i32]),
    "tg_table_load_pairs": (_i32, [_vp, _vp, _vp, _u64, _i32]),
    "tg_table_export": (_i32, [_vp, _u32, _u32, _i32, _i32, _pp, _pp, C.POINTER(_u64)]),
    "tg_histo": (_i32, [_vp, _vp]),
    "tg_cov_stats": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp, _vp, _vp, _vp]),
    "tg_label_bundles": (_i32, [_vp, _vp, _vp, _u64, _u32]),
    "tg_assign_reads": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp, _vp, _vp, _vp]),
    "tg_entropy_table": (None, [_i32, C.c_float, _vp]),
    "tg_dev_alloc": (_i32, [_vp, _u64, _pp]),
    "tg_dev_records_alloc": (_i32, [_vp, _u64, _pp]),
    "tg_dev_free": (_i32, [_vp, _vp]),
    "tg_memcpy_h2d": (_i32, [_vp, _vp, _vp, _u64]),
    "tg_memcpy_d2h": (_i32, [_vp, _vp, _vp, _u64]),
    "tg_memcpy_d2d": (_i32, [_vp, _vp, _vp, _u64]),
    "tg_memset_dev": (_i32, [_vp, _vp, _i32, _u64]),
    "tg_count_reads_dev": (_i32, [_vp, _vp, _u64, _i32]),
    "tg_count_partition_dev": (_i32, [_vp, _vp, _u64, _i32, _i32, _u32, _u32, _vp, _vp, _vp]),
    "tg_table_replay_log_dev": (_i32, [_vp, _vp, _vp, _vp, _u32, _u32]),
    "tg_log_refine_dev": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp, _u32, _u32, _u32, _u32]),
    "tg_log_entry_bytes": (_u32, []),
    "tg_table_set_count_floor": (_i32, [_vp, _u32]),
    "tg_query_partition_dev": (_i32, [_vp, _vp, _u64, _i32, _i32, _u32, _u32, _vp, _vp, _vp]),
    "tg_query_answer_dev": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp]),
    "tg_query_scatter_dev": (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _vp]),
    "tg_cov_stats_counts_dev": (_i32, [_vp, _vp, _vp, _u64, _i32, _u32, _vp, _vp, _vp, _vp]),
    "tg_count_records_dev": (_i32, [_vp, _vp, _vp, _u64, _i32]),
    "tg_records_pin_dev": (_i32, [_vp, _vp, _vp, _u64]),
    "tg_locus_prepare_dev": (_i32, [_vp, _i32, _i32]),
    "tg_weld_create": (_i32, [_vp, _i32, _vp, _u64, _pp]),
    "tg_weld_destroy": (None, [_vp]),
    "tg_weld_count_reads": (_i32, [_vp, _vp, _u64]),
    "tg_weld_count_reads_dev": (_i32, [_vp, _vp, _u64]),
    "tg_weld_counts": (_i32, [_vp, _vp]),
    "tg_table_count_sum": (_i32, [_vp, C.POINTER(_u64)]),
    "tg_valid_windows_dev": (_i32, [_vp, _vp, _u64, _i32, C.POINTER(_u64)]),
    "tg_records_hold": (_i32, [_vp, _vp, _u64]),
    "tg_records_release": (_i32, [_vp]),
    "tg_ipc_export": (_i32, [_vp, _vp, _vp]),
    "tg_ipc_open": (_i32, [_vp, _vp, _pp]),
    "tg_ipc_close": (_i32, [_vp, _vp]),
    "tg_count_partition_peers_dev": (_i32, [_vp, _vp, _u64, _i32, _i32, _u32, _u32, _u32, _u32, _pp, _vp, _vp]),
    "tg_cov_stats_dev": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp, _vp, _vp]),
    "tg_label_bundles_dev": (_i32, [_vp, _vp, _u64, _vp, _u64, _u32]),
    "tg_assign_reads_dev": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp, _vp, _vp]),
    "tg_timer_start": (_i32, [_vp]),
    "tg_timer_stop": (_i32, [_vp, C.POINTER(C.c_float)]),
    "tg_gups": (_i32, [_vp, _u64, _u64, _i32, _i32, C.POINTER(C.c_float)]),
    "tg_synth_reads_dev": (_i32, [_vp, _vp, _vp, _vp, _u32, _u64, _i32, _i32, _i32, _u32, _u32, _u64, _i32, _vp]),
}


class TrinityGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtrinity_gpu error {code}: {msg}")
        self.code = code


def load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C trinityrnaseq_b200`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(rc):
    if rc != TG_OK:
        raise TrinityGpuError(rc, lib().tg_last_error().decode("utf-8", "replace"))
