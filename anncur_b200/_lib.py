"""ctypes binding of libanncur_b200.so (include/anncur_b200.h).  There is no CPU fallback: if the
shared library is missing, or a call is made without a CUDA device, this raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libanncur_b200.so")

ABI_VERSION = 1
KIND_F32X3, KIND_BF16, KIND_F32R = 0, 1, 2
MAX_K, MAX_K_FUSED, MAX_K_DIM_F32R = 2048, 1024, 8192
E_INVALID, E_WORKSPACE, E_CUDA, E_UNSUPPORTED = -1, -2, -3, -4

_vp, _i, _i64, _sz, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double

# name -> (restype, argtypes); mirrors include/anncur_b200.h one to one
PROTOTYPES = {
    "anncur_abi_version": (_i, []),
    "anncur_last_error": (C.c_char_p, []),
    "anncur_pinv_workspace_bytes": (_sz, [_i, _i]),
    "anncur_pinv_f32": (_i, [_vp, _i, _i, _i, _d, _vp, _i, _vp, _vp, _sz, _vp]),
    "anncur_singular_values_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "anncur_orthonormalize_f32": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "anncur_jacobi_status": (_i, [_vp, _vp, _vp]),
    "anncur_gemm_f32": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "anncur_packed_items_bytes": (_sz, [_i64, _i, _i]),
    "anncur_pack_items": (_i, [_vp, _i64, _i64, _i, _i, _vp, _vp, _vp]),
    "anncur_score_topk_workspace_bytes": (_sz, [_i, _i64, _i, _i, _i]),
    "anncur_score_topk": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "anncur_score_dense_workspace_bytes": (_sz, [_i, _i64, _i, _i]),
    "anncur_score_dense": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp, _sz, _vp]),
    "anncur_score_bounds_dense": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp, _sz, _vp]),
    "anncur_recon_error_packed": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "anncur_score_topk_redo_rows": (_i, [_vp, _i, _i64, _i, _i, _i, C.POINTER(C.c_int), _vp]),
    "anncur_search_host_workspace_bytes": (_sz, [_i, _i64, _i, _i, _i]),
    "anncur_search_host": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "anncur_score_topk_f32_workspace_bytes": (_sz, [_i, _i64, _i, _i]),
    "anncur_score_topk_f32": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "anncur_topk_rows_f32": (_i, [_vp, _i64, _i, _i64, _i, _i64, _vp, _vp, _vp]),
    "anncur_merge_topk": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "anncur_topk_to_keys": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "anncur_merge_topk_keys_workspace_bytes": (_sz, [_i]),
    "anncur_merge_topk_keys": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "anncur_peer_channel_bytes": (_sz, [_i, _i, _i]),
    "anncur_peer_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "anncur_peer_free": (_i, [_vp]),
    "anncur_peer_export": (_i, [_vp, _vp]),
    "anncur_peer_open": (_i, [_vp, C.POINTER(_vp)]),
    "anncur_peer_close": (_i, [_vp]),
    "anncur_peer_scatter_keys": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, C.c_uint32, C.POINTER(_vp), _vp]),
    "anncur_peer_merge_owned": (_i, [_vp, _i, _i, _i, _i, _i, _i, C.c_uint32, _vp, _vp, _vp, _sz, _vp]),
    "anncur_peer_cert_failures": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_uint), _vp]),
    "anncur_peer_error": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_int), _vp]),
    "anncur_rerank_overlap": (_i, [_vp, _i64, _i, _i64, _vp, _i, _vp, _i, C.POINTER(C.c_int), _i, _vp, _vp, _vp, _vp]),
    "anncur_overlap_counts": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "anncur_recon_error_f32": (_i, [_vp, _i, _vp, _i64, _vp, _i64, _i, _i64, _i, _vp, _vp, _vp]),
    "anncur_adaptive_round_workspace_bytes": (_sz, [_i, _i, _i, _i64, _i]),
    "anncur_adaptive_round": (_i, [_vp, _i64, _i, _i64, _vp, _vp, _i, _i, _d, _i, _vp, _vp, _vp, _sz, _vp]),
    "anncur_adaptive_solve_workspace_bytes": (_sz, [_i, _i, _i, _i64]),
    "anncur_adaptive_solve": (_i, [_vp, _i64, _i, _i64, _vp, _vp, _vp, _i, _i, _d, _vp, _vp, _sz, _vp]),
    "anncur_transpose_f32": (_i, [_vp, _i64, _i, _i64, _vp, _vp]),
    "anncur_filter_excluded": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "anncur_score_topk_excluding_workspace_bytes": (_sz, [_i, _i64, _i, _i, _i, _i]),
    "anncur_score_topk_excluding": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _i, _i, _vp, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "anncur_adaptive_shared_bytes": (_sz, [_i, _i64, _i]),
    "anncur_adaptive_prepare_workspace_bytes": (_sz, [_i, _i64, _i]),
    "anncur_adaptive_prepare": (_i, [_vp, _i, _i64, _vp, _i, _d, _vp, _sz, _vp, _sz, _vp]),
    "anncur_adaptive_state_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "anncur_adaptive_begin": (_i, [_vp, _i, _i64, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "anncur_adaptive_extend": (_i, [_vp, _i, _i64, _vp, _i, _vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _sz, _vp]),
    "anncur_profile_enable": (_i, [_i]),
    "anncur_profile_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "anncur_kernel_launch_count": (C.c_uint64, []),
    "anncur_reset_kernel_launch_count": (None, []),
}


class AnncurError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libanncur_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python anncur_b200/csrc/build.py` "
            "(or __graft_entry__.build()).  anncur_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.anncur_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI version mismatch: library {lib.anncur_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise AnncurError(code, load().anncur_last_error().decode("utf-8", "replace"))
