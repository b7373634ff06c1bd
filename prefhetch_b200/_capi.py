"""ctypes binding of libprefhetch_b200.so (C ABI: include/prefhetch_b200.h).

The library is the product: if it is missing this module raises — there is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

# PF_LIB: an alternative build of the same library (kernel A/B experiments: python -m prefhetch_b200.build --variant NAME -D...)
LIB_PATH = Path(os.environ["PF_LIB"]) if os.environ.get("PF_LIB") else PKG / "libprefhetch_b200.so"

PF_MAX_PRIMES = 16
PF_T_COUNT = 8
PF_OK, PF_ERR_INVALID, PF_ERR_CUDA, PF_ERR_CAPACITY, PF_ERR_STATE, PF_ERR_FORMAT = range(6)
PHASES = {"coarse": 0, "to_ntt": 1, "rotate": 2, "mac": 3, "intt": 4}

# every symbol include/prefhetch_b200.h declares
EXPORTS = [
    "pf_abi_version", "pf_engine_create", "pf_engine_destroy", "pf_last_error", "pf_engine_stream",
    "pf_engine_set_stream", "pf_engine_synchronize", "pf_load_index", "pf_get_index_info", "pf_retrieve_centroids",
    "pf_coarse_quantize", "pf_search_lists_plain", "pf_load_pq", "pf_search_lists_pq", "pf_precise_search", "pf_set_galois_key", "pf_load_galois_keys",
    "pf_galois_elt_from_step", "pf_search_lists_encrypted", "pf_search_device", "pf_timing_enable", "pf_timing_read",
    "pf_launch_count", "pf_ipc_alloc", "pf_ipc_open", "pf_ipc_close", "pf_ipc_free", "pf_copy_async", "pf_flag_write", "pf_flag_wait", "pf_ntt_forward", "pf_ntt_inverse", "pf_ct_pt_mac", "pf_ct_add", "pf_ct_to_ntt",
    "pf_ct_from_ntt", "pf_rotate_rows", "pf_rotate_query_set", "pf_batch_encode", "pf_encode_block",
    "pf_ct_serialized_size", "pf_result_slot_size", "pf_result_serialized_size", "pf_set_result_parms_id", "pf_parms_id", "pf_seal_stream_inflate",
    "pf_ct_serialize", "pf_ct_deserialize", "pf_search_submit", "pf_search_collect", "pf_search_set_groups",
    "pf_host_register", "pf_host_unregister", "pf_device_checksum", "pf_seal_ct_expand", "pf_seal_ct_expand_device", "pf_seal_galois_keys_expand", "pf_seal_ct_expand_batch",
]


class PfParams(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("poly_degree", C.c_uint64),
                ("num_primes", C.c_uint32), ("dim", C.c_uint32), ("primes", C.c_uint64 * PF_MAX_PRIMES),
                ("plain_modulus", C.c_uint64), ("query_cts", C.c_uint32), ("partial_g", C.c_uint32),
                ("rank", C.c_uint32), ("world", C.c_uint32), ("result_limbs", C.c_uint32), ("reserved", C.c_uint32)]


class PfIndexInfo(C.Structure):
    _fields_ = [("nlist", C.c_uint64), ("ntotal", C.c_uint64), ("nblocks", C.c_uint64), ("nblocks_local", C.c_uint64),
                ("db_bytes", C.c_uint64), ("K", C.c_uint32), ("C", C.c_uint32), ("R", C.c_uint32),
                ("d_pad", C.c_uint32), ("L", C.c_uint32), ("k", C.c_uint32)]


class PfSearchStats(C.Structure):
    _fields_ = [("nresults", C.c_uint64), ("out_bytes", C.c_uint64), ("useful_distances", C.c_uint64),
                ("slot_distances", C.c_uint64)]


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m prefhetch_b200.build` "
                           "(the CUDA extension is the product; there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp, u64p, i64p, f32p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_float), \
        C.POINTER(C.c_uint8)
    i32p, szp = C.POINTER(C.c_int32), C.POINTER(C.c_size_t)
    sig = {
        "pf_abi_version": ([], C.c_int),
        "pf_engine_create": ([C.POINTER(PfParams), C.POINTER(vp)], C.c_int),
        "pf_engine_destroy": ([vp], None),
        "pf_last_error": ([vp], C.c_char_p),
        "pf_engine_stream": ([vp], vp),
        "pf_engine_set_stream": ([vp, vp], C.c_int),
        "pf_engine_synchronize": ([vp], C.c_int),
        "pf_load_index": ([vp, C.c_uint64, f32p, i64p, i64p, f32p], C.c_int),
        "pf_get_index_info": ([vp, C.POINTER(PfIndexInfo)], C.c_int),
        "pf_retrieve_centroids": ([vp, f32p, C.c_uint64], C.c_int),
        "pf_coarse_quantize": ([vp, C.c_uint64, f32p, C.c_uint32, i64p, f32p], C.c_int),
        "pf_search_lists_plain": ([vp, C.c_uint64, f32p, i64p, C.c_uint32, f32p, i64p, C.c_uint64, u64p, u64p], C.c_int),
        "pf_load_pq": ([vp, C.c_uint32, C.c_uint32, f32p, C.POINTER(C.c_uint8)], C.c_int),
        "pf_search_lists_pq": ([vp, C.c_uint64, f32p, i64p, C.c_uint32, f32p, i64p, C.c_uint64, u64p, u64p], C.c_int),
        "pf_precise_search": ([vp, C.c_uint64, f32p, i64p, C.c_uint32, f32p], C.c_int),
        "pf_set_galois_key": ([vp, C.c_uint32, u64p], C.c_int),
        "pf_load_galois_keys": ([vp, u8p, C.c_size_t], C.c_int),
        "pf_galois_elt_from_step": ([vp, C.c_int], C.c_uint32),
        "pf_search_lists_encrypted": ([vp, C.c_uint64, vp, C.c_uint64, u64p, i64p, C.c_uint32, vp, C.c_uint64, u64p,
                                       C.c_uint64, u64p, i64p, C.c_uint64, u64p, u64p, C.POINTER(PfSearchStats)], C.c_int),
        "pf_search_submit": ([vp, C.c_uint64, vp, C.c_uint64, u64p, i64p, C.c_uint32, vp, C.c_uint64, u64p, C.c_uint64,
                              u64p, i64p, C.c_uint64, u64p, u64p, C.POINTER(PfSearchStats), u64p], C.c_int),
        "pf_search_collect": ([vp, C.c_uint64], C.c_int),
        "pf_search_set_groups": ([vp, C.c_uint32], C.c_int),
        "pf_host_register": ([vp, vp, C.c_size_t], C.c_int),
        "pf_host_unregister": ([vp, vp], C.c_int),
        "pf_device_checksum": ([vp, vp, C.c_uint64, u64p, vp], C.c_int),
        "pf_search_device": ([vp, C.c_uint64, vp, i64p, C.c_uint32, vp, C.c_uint64, u64p, C.POINTER(PfSearchStats)],
                             C.c_int),
        "pf_timing_enable": ([vp, C.c_int], C.c_int),
        "pf_timing_read": ([vp, f32p, u64p, C.c_int], C.c_int),
        "pf_launch_count": ([vp], C.c_uint64),
        "pf_ipc_alloc": ([vp, C.c_size_t, C.POINTER(vp), u8p], C.c_int),
        "pf_ipc_open": ([vp, u8p, C.POINTER(vp)], C.c_int),
        "pf_ipc_close": ([vp, vp], C.c_int),
        "pf_ipc_free": ([vp, vp], C.c_int),
        "pf_copy_async": ([vp, vp, vp, C.c_size_t, vp], C.c_int),
        "pf_flag_write": ([vp, vp, C.c_uint32, vp], C.c_int),
        "pf_flag_wait": ([vp, vp, C.c_uint32, vp], C.c_int),
        "pf_ntt_forward": ([vp, u64p, C.c_uint64, i32p], C.c_int),
        "pf_ntt_inverse": ([vp, u64p, C.c_uint64, i32p], C.c_int),
        "pf_ct_pt_mac": ([vp, u64p, u64p, C.c_uint32, u64p, u64p], C.c_int),
        "pf_ct_add": ([vp, u64p, u64p, u64p], C.c_int),
        "pf_ct_to_ntt": ([vp, u64p, C.c_uint64], C.c_int),
        "pf_ct_from_ntt": ([vp, u64p, C.c_uint64], C.c_int),
        "pf_rotate_rows": ([vp, u64p, C.c_int, u64p], C.c_int),
        "pf_rotate_query_set": ([vp, u64p, C.c_int, u64p], C.c_int),
        "pf_batch_encode": ([vp, u64p, u64p], C.c_int),
        "pf_encode_block": ([vp, i32p, C.c_uint32, u64p, u64p], C.c_int),
        "pf_ct_serialized_size": ([vp], C.c_size_t),
        "pf_result_slot_size": ([vp], C.c_size_t),
        "pf_result_serialized_size": ([vp], C.c_size_t),
        "pf_set_result_parms_id": ([vp, u64p], C.c_int),
        "pf_parms_id": ([C.c_uint64, u64p, C.c_uint32, C.c_uint64, u64p], C.c_int),
        "pf_seal_stream_inflate": ([vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)], C.c_int),
        "pf_seal_ct_expand": ([vp, C.c_size_t, C.c_uint64, u64p, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_size_t),
                               C.POINTER(C.c_size_t)], C.c_int),
        "pf_seal_ct_expand_device": ([vp, vp, C.c_size_t, u64p, C.c_size_t], C.c_int),
        "pf_seal_ct_expand_batch": ([vp, C.c_size_t, u64p, C.c_uint64, C.c_uint64, u64p, C.c_uint32, vp, C.c_size_t, u64p, C.c_uint32], C.c_int),
        "pf_seal_galois_keys_expand": ([vp, C.c_size_t, C.c_uint64, u64p, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_size_t)], C.c_int),
        "pf_ct_serialize": ([vp, u64p, C.c_int, u8p, C.c_size_t, szp], C.c_int),
        "pf_ct_deserialize": ([vp, u8p, C.c_size_t, u64p, C.c_size_t, i32p, i32p, szp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    _lib = lib
    return lib
