"""ctypes binding of libnann_b200.so (the C ABI in include/nann_b200.h).

There is deliberately no fallback: if the shared library is missing, or no CUDA device is
present, the compute entry points raise.  Nothing here imports or calls oracle/.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnann_b200.so")

OK, INVALID_ARGUMENT, DEADLINE_EXCEEDED, NOT_FOUND = 0, 3, 4, 5
RESOURCE_EXHAUSTED, FAILED_PRECONDITION, UNIMPLEMENTED, INTERNAL = 8, 9, 12, 13
F16, F32, F64, I32, I64 = 0, 1, 2, 3, 4
SCORER_EXACT, SCORER_TENSOR = 0, 1

_CODE_NAME = {3: "InvalidArgument", 4: "DeadlineExceeded", 5: "NotFound", 8: "ResourceExhausted",
              9: "FailedPrecondition", 12: "Unimplemented", 13: "Internal"}


class NannError(RuntimeError):
    """Mirror of a non-OK tensorflow::Status raised by an op kernel."""

    def __init__(self, code, message):
        super().__init__(f"{_CODE_NAME.get(code, code)}: {message}")
        self.code = code
        self.message = message


ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_int64)


class SearchStats(C.Structure):
    _fields_ = [("n_scored", C.c_int64 * 5), ("n_expanded", C.c_int64 * 5), ("n_failed", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m nann_b200.build` "
            "(nvcc, sm_100a). nann_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.nann_abi_version.restype = C.c_int
    L.nann_last_error.restype = C.c_char_p
    L.nann_kernel_launch_count.restype = C.c_uint64
    L.nann_device_info.argtypes = [i32] + [vp] * 5
    L.nann_npy_peek.argtypes = [C.c_char_p, vp, vp, vp]
    L.nann_huge_const_create.argtypes = [C.c_char_p, i32, vp, i32, i32, vp]
    for f in ("nann_huge_const_host", "nann_huge_const_device"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    L.nann_huge_const_bytes.restype = i64
    L.nann_huge_const_bytes.argtypes = [vp]
    L.nann_huge_const_destroy.restype = None
    L.nann_huge_const_destroy.argtypes = [vp]
    for f in ("nann_group_gather_i32", "nann_group_gather_i64"):
        getattr(L, f).argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i32, ALLOC_FN, vp, vp]
    for f in ("nann_bitmap_ref_difference_i32", "nann_bitmap_ref_difference_i64"):
        getattr(L, f).argtypes = [vp, i64, vp, i64, vp, i64, ALLOC_FN, vp, vp]
    for f in ("nann_bloom_filter_difference_i32", "nann_bloom_filter_difference_i64"):
        getattr(L, f).argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, ALLOC_FN, vp, vp]
    L.nann_topk_v2_f32.argtypes = [vp, i64, i64, C.c_int32, i32, vp, vp, vp]
    L.nann_gather_rows.argtypes = [vp, i64, i64, vp, i64, vp, vp]
    L.nann_scorer_create_mlp.argtypes = [i32, i32, vp, vp, vp, vp, vp, i32, vp]
    L.nann_scorer_create_attention.argtypes = [vp, i64, i32, vp]
    L.nann_scorer_attention_blob_size.restype = i64
    L.nann_scorer_set_precision.argtypes = [vp, i32]
    L.nann_scorer_user_floats.argtypes = [vp]
    L.nann_scorer_item_dim.argtypes = [vp]
    L.nann_scorer_set_admission.argtypes = [vp, i32, i32, i32]
    L.nann_scorer_admission_state.argtypes = [vp, vp, vp, vp, vp, vp]
    L.nann_scorer_destroy.restype = None
    L.nann_scorer_destroy.argtypes = [vp]
    L.nann_blaze_xla_run.argtypes = [vp, vp, vp, i64, vp, vp]
    L.nann_scorer_run_ids.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp]
    L.nann_index_create.argtypes = [i64, i32, vp, i32, vp, vp, i32, i64, vp, vp, i32, vp, i32, vp]
    L.nann_index_load.argtypes = [C.c_char_p, C.c_char_p, i32, vp]
    L.nann_index_n_items.restype = i64
    L.nann_index_n_items.argtypes = [vp]
    L.nann_index_dim.argtypes = [vp]
    L.nann_index_n_enter_points.restype = i64
    L.nann_index_n_enter_points.argtypes = [vp]
    L.nann_index_emb_device.restype = vp
    L.nann_index_emb_device.argtypes = [vp]
    L.nann_index_destroy.restype = None
    L.nann_index_destroy.argtypes = [vp]
    L.nann_searcher_create.argtypes = [vp, vp, i32, vp, vp]
    L.nann_searcher_destroy.restype = None
    L.nann_searcher_destroy.argtypes = [vp]
    L.nann_searcher_set_trace.argtypes = [vp, i32]
    L.nann_search_batch.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.nann_searcher_get_trace.argtypes = [vp, i32, i32, vp, vp, i64, vp]
    L.nann_searcher_get_nodes.argtypes = [vp, vp, i64]
    L.nann_searcher_set_profile.argtypes = [vp, i32]
    L.nann_searcher_get_profile.argtypes = [vp, vp, vp, vp, vp]
    L.nann_merge_topk.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp]
    L.nann_shard_group_create.argtypes = [i32, i32, i32, i32, i32, vp]
    L.nann_shard_group_export.argtypes = [vp, vp]
    L.nann_shard_group_connect.argtypes = [vp, vp]
    L.nann_shard_group_connect_local.argtypes = [vp, i32]
    L.nann_shard_group_destroy.restype = None
    L.nann_shard_group_destroy.argtypes = [vp]
    L.nann_search_sharded.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, vp, vp]
    L.nann_shard_group_wait.argtypes = [vp, vp, i32]
    L.nann_search_sharded_push.argtypes = [vp, vp, vp, i32, vp, vp]
    L.nann_search_sharded_merge.argtypes = [vp, i32, vp, vp, vp]
    L.nann_index_create_sharded.argtypes = [i64, i32, vp, i32, i64, i64, vp, vp, i32, i64, vp, vp, i32, vp, i32, vp]
    L.nann_dist_group_create.argtypes = [vp, i32, i32, vp]
    L.nann_dist_group_export.argtypes = [vp, vp]
    L.nann_dist_group_connect.argtypes = [vp, vp]
    L.nann_dist_group_connect_local.argtypes = [vp, i32]
    L.nann_dist_group_check.argtypes = [vp]
    L.nann_dist_group_destroy.restype = None
    L.nann_dist_group_destroy.argtypes = [vp]
    L.nann_search_distributed.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.nann_eval_searcher_create.argtypes = [vp, vp, i32, vp, i32, vp]
    L.nann_eval_searcher_destroy.restype = None
    L.nann_eval_searcher_destroy.argtypes = [vp]
    L.nann_search_eval_batch.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


def check(status):
    if status != OK:
        raise NannError(status, lib().nann_last_error().decode(errors="replace"))


def launch_count():
    return int(lib().nann_kernel_launch_count())


def device_info(device=0):
    n, sm, cc_a, cc_b = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    hbm = C.c_int64(0)
    check(lib().nann_device_info(device, C.byref(n), C.byref(sm), C.byref(hbm), C.byref(cc_a), C.byref(cc_b)))
    return dict(device_count=n.value, sm_count=sm.value, hbm_bytes=hbm.value, cc=(cc_a.value, cc_b.value))
