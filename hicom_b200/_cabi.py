"""ctypes binding of include/hicom_b200.h.  No fallback: a missing library is an ImportError-class failure."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhicom_b200.so")

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_GELU_TANH = 0, 1, 2
Q_POOLED, Q_FILM_LN, Q_VECTOR, Q_EXPLICIT = 0, 1, 2, 3
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2

# name -> (restype, argtypes); mirrors include/hicom_b200.h one to one
PROTOTYPES = {
    "hicom_abi_version": (c_int, []),
    "hicom_last_error": (c_char_p, []),
    "hicom_kernel_launch_count": (ctypes.c_uint64, []),
    "hicom_kernel_timing_enable": (c_int, [c_int]),
    "hicom_kernel_timing_collect": (c_size_t, [ctypes.c_char_p, c_size_t]),
    "hicom_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "hicom_set_sm_limit": (c_int, [c_int]),
    "hicom_grid_pool": (c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "hicom_local_attend": (c_int, [c_void_p] * 8 + [c_int] * 8 + [c_float, c_int, c_int, c_void_p]),
    "hicom_linear": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                             c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_int, c_void_p]),
    "hicom_layernorm": (c_int, [c_void_p] * 4 + [c_int64, c_int, c_int, c_void_p]),
    "hicom_film_layernorm": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "hicom_add_layernorm": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    "hicom_mix_layernorm": (c_int, [c_void_p] * 6 + [c_int64, c_int, c_int, c_void_p]),
    "hicom_guide_attend": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_float, c_int, c_void_p]),
    "hicom_global_fold_query": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "hicom_global_attend_workspace_bytes": (c_size_t, [c_int] * 9),
    "hicom_global_attend_partial": (c_int, [c_void_p] * 8 + [c_int] * 8 + [c_void_p, c_size_t, c_int, c_void_p]),
    "hicom_global_attend_partial_keys": (c_int, [c_void_p] * 9 + [c_int] * 8 + [c_void_p, c_size_t, c_int, c_void_p]),
    "hicom_posadd": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "hicom_l2norm_rows": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p]),
    "hicom_softmax_merge": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p, c_int, c_void_p]),
    "hicom_softmax_merge_lse": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p, c_int, c_void_p, c_void_p]),
    "hicom_shard_combine": (c_int, [c_void_p] + [ctypes.c_longlong] * 3 + [c_int] * 5 + [c_void_p, c_int, c_void_p]),
    "hicom_softmax_reduce": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p] * 4),
    "hicom_global_value_proj": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "hicom_gemm": (c_int, [c_void_p] + [c_int64] * 4 + [c_void_p] + [c_int64] * 4 + [c_void_p] + [c_int64] * 3 +
                   [c_int] * 5 + [c_float] + [c_int] * 3 + [c_void_p]),
    "hicom_colsum": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "hicom_act_backward": (c_int, [c_void_p] * 3 + [c_int64, c_int, c_int, c_int, c_void_p]),
    "hicom_col_stats": (c_int, [c_void_p] * 3 + [c_int, ctypes.c_longlong, c_int, c_int, c_void_p]),
    "hicom_softmax_backward": (c_int, [c_void_p] * 5 + [c_int, c_int64, c_int, c_int, c_void_p]),
    "hicom_grid_pool_backward": (c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "hicom_l2norm_rows_backward": (c_int, [c_void_p] * 3 + [ctypes.c_longlong, c_int, c_int, c_void_p]),
    "hicom_local_attend_backward": (c_int, [c_void_p] * 7 + [c_int] * 7 + [c_float, c_int, c_int, c_void_p]),
    "hicom_film_layernorm_backward": (c_int, [c_void_p] * 8 + [c_int64, c_int, c_int, c_int, c_void_p]),
    "hicom_mix_layernorm_backward": (c_int, [c_void_p] * 11 + [c_int64, c_int, c_int, c_void_p]),
}

_lib = None


class HicomLibraryError(RuntimeError):
    pass


def load():
    """Load libhicom_b200.so (built by ``python -m hicom_b200.build``).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HicomLibraryError(
            f"{LIB_PATH} not found: build it with `python -m hicom_b200.build` "
            "(hicom_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    got = lib.hicom_abi_version()
    if got != 1:
        raise HicomLibraryError(f"ABI version mismatch: library {got}, binding 1")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().hicom_last_error()
        raise RuntimeError(f"{what}: {msg.decode() if msg else 'unknown error'}")
