"""ctypes binding of include/b200ann.h -- the same C ABI a JVM would bind through JNI.

Fails loudly when libb200ann.so is missing: there is no CPU fallback behind this module.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

_HERE = Path(__file__).resolve().parent
import os

# B200ANN_LIB lets a developer A/B two builds of the same ABI (point it at the other libb200ann.so)
LIB_PATH = Path(os.environ["B200ANN_LIB"]) if os.environ.get("B200ANN_LIB") else _HERE / "lib" / "libb200ann.so"

ANN_OK = 0
ANN_ERR_INVALID_ARGUMENT = -1
ANN_ERR_NULL_POINTER = -2
ANN_ERR_DIMENSION_MISMATCH = -3
ANN_ERR_NEGATIVE_K = -4
ANN_ERR_NO_DEVICE = -5
ANN_ERR_CUDA = -6
ANN_ERR_OUT_OF_MEMORY = -7
ANN_ERR_CANDIDATE_OVERFLOW = -8
ANN_ERR_UNKNOWN_OPTION = -9

ANN_FLAG_L2_SQUARED = 0x1
ANN_FLAG_NO_SHADOW = 0x2
ANN_FLAG_ACCUM_F32 = 0x4
ANN_FLAG_COSINE_UNIT_ROWS = 0x8
ANN_ID_AUTO, ANN_ID_INT64_BE, ANN_ID_INT32_BE = 0, 1, 2
ANN_LAYOUT_FLOAT_TENSOR, ANN_LAYOUT_DOUBLE_TENSOR, ANN_LAYOUT_RAW_FLOAT = 0, 1, 2

# every symbol include/b200ann.h declares (tests/test_capi_symbols.py checks header <-> library <-> this list)
SYMBOLS = (
    "ann_create", "ann_destroy", "ann_append_batch", "ann_append_batch_device", "ann_update_batch", "ann_read_rows", "ann_size", "ann_query_batch",
    "ann_query_batch_device", "ann_merge_topk_device", "ann_exchange_merge_device", "ann_result_block_bytes", "ann_query_seed_device", "ann_query_finish_device",
    "ann_query_filter_device", "ann_query_rescore_device", "ann_exchange_merge_slice_device",
    "ann_query_seed_push_device", "ann_query_filter_push_device", "ann_peer_push_device",
    "ann_query_seed_slice_push_device", "ann_query_filter_bounds_push_device",
    "ann_sharded_create", "ann_sharded_destroy", "ann_sharded_append_batch", "ann_sharded_size", "ann_sharded_query_batch",
    "ann_sharded_shard", "ann_sharded_set_option", "ann_sharded_get_stat",
    "ann_save_directory", "ann_load_directory", "ann_sharded_save_directory", "ann_sharded_load_directory",
    "ann_persisted_embedding_encode", "ann_persisted_embedding_decode", "ann_loadtest", "ann_knn_join", "ann_distance_pairs", "ann_normalize_rows", "ann_set_option", "ann_get_stat", "ann_last_error", "ann_version",
)


class AnnConfig(ctypes.Structure):
    _fields_ = [("metric", ctypes.c_int32), ("dim", ctypes.c_int32), ("capacity_hint", ctypes.c_int64),
                ("device", ctypes.c_int32), ("flags", ctypes.c_uint32)]


class AnnLoadStats(ctypes.Structure):
    _fields_ = [("qps", ctypes.c_double), ("avg_us", ctypes.c_double), ("p50_us", ctypes.c_double), ("p90_us", ctypes.c_double),
                ("p99_us", ctypes.c_double), ("wall_seconds", ctypes.c_double), ("calls", ctypes.c_int64),
                ("device_batches", ctypes.c_int64), ("mismatches", ctypes.c_int64)]


class AnnError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b200ann error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python the-algorithm_b200/build.py` "
                "(there is no CPU fallback for the CUDA engine)")
        L = ctypes.CDLL(str(LIB_PATH))
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.ann_create.restype = ctypes.c_int
        L.ann_create.argtypes = [ctypes.POINTER(AnnConfig), ctypes.POINTER(vp)]
        L.ann_destroy.restype = None
        L.ann_destroy.argtypes = [vp]
        L.ann_append_batch.restype = ctypes.c_int
        L.ann_append_batch.argtypes = [vp, vp, vp, i64]
        L.ann_append_batch_device.restype = ctypes.c_int
        L.ann_append_batch_device.argtypes = [vp, vp, vp, i64, vp]
        L.ann_update_batch.restype = ctypes.c_int
        L.ann_update_batch.argtypes = [vp, vp, vp, i64]
        L.ann_read_rows.restype = ctypes.c_int
        L.ann_read_rows.argtypes = [vp, i64, i64, vp, vp]
        L.ann_size.restype = ctypes.c_int
        L.ann_size.argtypes = [vp, ctypes.POINTER(i64)]
        L.ann_query_batch.restype = ctypes.c_int
        L.ann_query_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
        L.ann_query_batch_device.restype = ctypes.c_int
        L.ann_query_batch_device.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
        L.ann_merge_topk_device.restype = ctypes.c_int
        L.ann_merge_topk_device.argtypes = [i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
        L.ann_exchange_merge_device.restype = ctypes.c_int
        L.ann_exchange_merge_device.argtypes = [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), i32, i32, i32, i32, i32, vp]
        L.ann_query_seed_device.restype = ctypes.c_int
        L.ann_query_seed_device.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        L.ann_query_finish_device.restype = ctypes.c_int
        L.ann_query_finish_device.argtypes = [vp, vp, i32, i32, i32, ctypes.POINTER(vp), i32, vp, vp, vp, vp]
        L.ann_query_filter_device.restype = ctypes.c_int
        L.ann_query_filter_device.argtypes = [vp, vp, i32, i32, i32, ctypes.POINTER(vp), i32, vp, vp]
        L.ann_query_seed_push_device.restype = ctypes.c_int
        L.ann_query_seed_push_device.argtypes = [vp, vp, i32, i32, i32, ctypes.POINTER(vp), i32, vp]
        L.ann_query_filter_push_device.restype = ctypes.c_int
        L.ann_query_filter_push_device.argtypes = [vp, vp, i32, i32, i32, ctypes.POINTER(vp), i32, ctypes.POINTER(vp), i32, vp]
        L.ann_query_seed_slice_push_device.restype = ctypes.c_int
        L.ann_query_seed_slice_push_device.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, ctypes.POINTER(vp), i32, vp]
        L.ann_query_filter_bounds_push_device.restype = ctypes.c_int
        L.ann_query_filter_bounds_push_device.argtypes = [vp, vp, i32, i32, i32, vp, i32, ctypes.POINTER(vp), i32, vp]
        L.ann_peer_push_device.restype = ctypes.c_int
        L.ann_peer_push_device.argtypes = [i32, vp, ctypes.POINTER(vp), i32, ctypes.c_size_t, vp]
        L.ann_query_rescore_device.restype = ctypes.c_int
        L.ann_query_rescore_device.argtypes = [vp, vp, i32, i32, i32, ctypes.POINTER(vp), i32, vp, vp, vp, vp]
        L.ann_exchange_merge_slice_device.restype = ctypes.c_int
        L.ann_exchange_merge_slice_device.argtypes = [i32, ctypes.POINTER(vp), i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.ann_sharded_create.restype = ctypes.c_int
        L.ann_sharded_create.argtypes = [ctypes.POINTER(AnnConfig), ctypes.POINTER(i32), i32, ctypes.POINTER(vp)]
        L.ann_sharded_destroy.restype = None
        L.ann_sharded_destroy.argtypes = [vp]
        L.ann_sharded_append_batch.restype = ctypes.c_int
        L.ann_sharded_append_batch.argtypes = [vp, vp, vp, i64]
        L.ann_sharded_size.restype = ctypes.c_int
        L.ann_sharded_size.argtypes = [vp, ctypes.POINTER(i64)]
        L.ann_sharded_query_batch.restype = ctypes.c_int
        L.ann_sharded_query_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
        L.ann_sharded_shard.restype = ctypes.c_int
        L.ann_sharded_shard.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(i64)]
        L.ann_sharded_set_option.restype = ctypes.c_int
        L.ann_sharded_set_option.argtypes = [vp, ctypes.c_char_p, i64]
        L.ann_sharded_get_stat.restype = ctypes.c_int
        L.ann_sharded_get_stat.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(i64)]
        L.ann_save_directory.restype = ctypes.c_int
        L.ann_save_directory.argtypes = [vp, ctypes.c_char_p, i32, i32]
        L.ann_load_directory.restype = ctypes.c_int
        L.ann_load_directory.argtypes = [ctypes.POINTER(AnnConfig), ctypes.c_char_p, i32, ctypes.POINTER(vp)]
        L.ann_sharded_save_directory.restype = ctypes.c_int
        L.ann_sharded_save_directory.argtypes = [vp, ctypes.c_char_p, i32, i32]
        L.ann_sharded_load_directory.restype = ctypes.c_int
        L.ann_sharded_load_directory.argtypes = [ctypes.POINTER(AnnConfig), ctypes.c_char_p, i32, ctypes.POINTER(i32), i32, ctypes.POINTER(vp)]
        L.ann_persisted_embedding_encode.restype = ctypes.c_int64
        L.ann_persisted_embedding_encode.argtypes = [i64, i32, vp, i32, i32, vp, i64]
        L.ann_persisted_embedding_decode.restype = ctypes.c_int
        L.ann_persisted_embedding_decode.argtypes = [vp, i64, i32, ctypes.POINTER(i64), vp, i32, ctypes.POINTER(i32), ctypes.POINTER(i64)]
        L.ann_loadtest.restype = ctypes.c_int
        L.ann_loadtest.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, ctypes.POINTER(AnnLoadStats)]
        L.ann_result_block_bytes.restype = ctypes.c_size_t
        L.ann_result_block_bytes.argtypes = [i32, i32]
        L.ann_knn_join.restype = ctypes.c_int
        L.ann_knn_join.argtypes = [ctypes.POINTER(AnnConfig), vp, vp, i64, vp, i64, i32, i64, i32, vp, vp, vp]
        L.ann_distance_pairs.restype = ctypes.c_int
        L.ann_distance_pairs.argtypes = [i32, ctypes.c_uint32, i32, vp, vp, i64, vp, i32]
        L.ann_normalize_rows.restype = ctypes.c_int
        L.ann_normalize_rows.argtypes = [i32, vp, i64, vp, i32]
        L.ann_set_option.restype = ctypes.c_int
        L.ann_set_option.argtypes = [vp, ctypes.c_char_p, i64]
        L.ann_get_stat.restype = ctypes.c_int
        L.ann_get_stat.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(i64)]
        L.ann_last_error.restype = ctypes.c_char_p
        L.ann_last_error.argtypes = []
        L.ann_version.restype = ctypes.c_int
        L.ann_version.argtypes = []
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != ANN_OK:
        raise AnnError(rc, lib().ann_last_error().decode("utf-8", "replace"))
