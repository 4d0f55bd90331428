"""ctypes view of include/bgn_b200.h.

Loads bgn_b200/libbgn_b200.so (built in-tree by `make` / __graft_entry__.build()).
There is no fallback of any kind: if the CUDA library is missing or a symbol
does not resolve, importing a compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BGN_B200_LIB") or os.path.join(HERE, "libbgn_b200.so")  # env: developer A/B builds

BGN_OK, BGN_E_BADARG, BGN_E_CUDA, BGN_E_NOTSETUP, BGN_E_NOMEM, BGN_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5

u8p = C.c_void_p  # data pointers are passed as raw addresses (host or device)


class bgn_params(C.Structure):
    _fields_ = [("p_be", C.c_char_p), ("p_len", C.c_size_t), ("n_be", C.c_char_p), ("n_len", C.c_size_t),
                ("l", C.c_uint64), ("P_bytes", C.c_char_p), ("Q_bytes", C.c_char_p)]


# name -> (restype, argtypes); must list every symbol include/bgn_b200.h declares
SIGNATURES = {
    "bgn_ctx_create": (C.c_int, [C.POINTER(bgn_params), C.c_int, C.POINTER(C.c_void_p)]),
    "bgn_ctx_destroy": (None, [C.c_void_p]),
    "bgn_last_error": (C.c_char_p, [C.c_void_p]),
    "bgn_global_last_error": (C.c_char_p, []),
    "bgn_ctx_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bgn_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_long]),
    "bgn_ctx_set_secret": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint32]),
    "bgn_encrypt_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_g1_add_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_g1_sub_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_g1_neg_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, u8p]),
    "bgn_g1_mulconst_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, C.c_size_t, u8p]),
    "bgn_gt_mul_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_gt_div_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_gt_inv_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, u8p]),
    "bgn_gt_pow_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, C.c_size_t, u8p]),
    "bgn_pair_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_make_l2_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, u8p]),
    "bgn_multpoly_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, u8p, C.c_size_t, C.c_size_t, u8p]),
    "bgn_l2_sum_reduce": (C.c_int, [C.c_void_p, u8p, C.c_size_t, C.c_size_t, u8p]),
    "bgn_gt_pow_secret_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, u8p]),
    "bgn_decrypt_batch": (C.c_int, [C.c_void_p, u8p, C.c_int, C.c_size_t, u8p, u8p]),
    "bgn_g1_blind_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_gt_blind_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_multconstpoly_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, C.c_int,
                                           C.c_size_t, u8p]),
    "bgn_evalpoly_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, C.c_int, C.c_uint32, C.c_size_t, u8p]),
    "bgn_make_poly_l2_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, C.c_size_t, u8p]),
    "bgn_buf_import": (C.c_int, [C.c_void_p, C.c_int, u8p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "bgn_buf_export": (C.c_int, [C.c_void_p, C.c_void_p, u8p]),
    "bgn_buf_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "bgn_buf_free": (None, [C.c_void_p]),
    "bgn_encrypt_h": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "bgn_g1_add_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bgn_gt_mul_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bgn_pair_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "bgn_multpoly_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                 C.POINTER(C.c_void_p)]),
    "bgn_l2_sum_reduce_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]),
    "bgn_decrypt_h": (C.c_int, [C.c_void_p, C.c_void_p, u8p, u8p]),
    "bgn_group_create": (C.c_int, [C.POINTER(bgn_params), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "bgn_group_destroy": (None, [C.c_void_p]),
    "bgn_group_size": (C.c_int, [C.c_void_p]),
    "bgn_group_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "bgn_group_last_error": (C.c_char_p, [C.c_void_p]),
    "bgn_group_set_secret": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint32]),
    "bgn_group_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_long]),
    "bgn_group_encrypt_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_group_g1_add_batch": (C.c_int, [C.c_void_p, u8p, u8p, C.c_size_t, u8p]),
    "bgn_group_multpoly_batch": (C.c_int, [C.c_void_p, u8p, C.c_size_t, u8p, C.c_size_t, C.c_size_t, u8p]),
    "bgn_group_decrypt_batch": (C.c_int, [C.c_void_p, u8p, C.c_int, C.c_size_t, u8p, u8p]),
    "bgn_group_inner_product": (C.c_int, [C.c_void_p, u8p, C.c_size_t, u8p, C.c_size_t, C.c_size_t, u8p]),
    "bgn_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "bgn_timing_reset": (C.c_int, [C.c_void_p]),
    "bgn_timing_get": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "bgn_timing_last_call": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "bgn_bench_mulmod": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "bgn_bench_issue_mix": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                      C.POINTER(C.c_double)]),
    "bgn_bench_imad_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                      C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the CUDA library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "bgn_b200: %s is missing -- build it with `make` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
