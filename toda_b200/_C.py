"""ctypes binding of libtoda_b200.so (C ABI declared in include/toda_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtoda_b200.so")

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
_IP = ctypes.POINTER(ctypes.c_int)
_FP = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); must list every symbol include/toda_b200.h declares
SIGNATURES = {
    "toda_last_error": (ctypes.c_char_p, []),
    "toda_version": (c_int, []),
    "toda_device_info": (c_int, [_IP, _IP, _IP]),
    "toda_index_bytes": (c_sz, [c_int] * 4),
    "toda_index_insert": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "toda_index_insert_strided": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, _IP, _IP, _IP, c_vp]),
    "toda_index_build": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp]),
    "toda_index_rows": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp]),
    "toda_index_release": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "toda_points_select_workspace_bytes": (c_sz, [c_int]),
    "toda_points_select": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_int, ctypes.POINTER(ctypes.c_double), c_int, c_int,
                                   c_vp, c_vp, c_vp, c_sz, c_vp]),
    "toda_voxelize_workspace_bytes": (c_sz, [c_i64, c_int, _IP, c_int, c_int]),
    "toda_voxelize_hard": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_int, _FP, _FP, _IP, c_int, c_int,
                                   c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "toda_mean_vfe_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "toda_mean_vfe_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "toda_rulebook_subm": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, _IP, c_vp, c_vp]),
    "toda_rulebook_sparse": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp,
                                     c_int, _IP, _IP, _IP, c_vp, c_vp, c_vp]),
    "toda_weight_repack": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "toda_spconv_fwd_workspace_bytes": (c_sz, [c_int] * 5),
    "toda_spconv_uses_tensor_cores": (c_int, [c_int] * 4),
    "toda_parity_order_workspace_bytes": (c_sz, [c_int]),
    "toda_parity_order": (c_int, [c_vp, c_int, _IP, _IP, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "toda_table_tile_masks": (c_int, [c_vp, c_int, c_int, c_vp, c_vp]),
    "toda_spconv_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp,
                                c_sz, c_vp]),
    "toda_tile_plan_capacity": (c_int, [c_int]),
    "toda_table_tile_plan": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "toda_spconv_fwd_plan": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_sz, c_vp]),
    "toda_spconv_wgrad_workspace_bytes": (c_sz, [c_int] * 6),
    "toda_spconv_wgrad": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_sz, c_int, c_vp]),
    "toda_bn_workspace_bytes": (c_sz, [c_int]),
    "toda_bn_stats": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                              c_sz, c_vp]),
    "toda_bn_finalize_sums": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "toda_bn_eval_coeffs": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_int, c_vp, c_vp, c_vp]),
    "toda_bn_apply": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "toda_bn_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                            c_vp, c_sz, c_vp]),
    "toda_col_sum": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    "toda_bev_scatter_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "toda_bev_scatter_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "toda_gather_rows": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
}

_lib = None


def lib():
    """The loaded shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  toda_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().toda_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libtoda_b200 {what} failed (code {rc}): {msg}")


def ints(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def doubles(vals):
    return (ctypes.c_double * len(vals))(*[float(v) for v in vals])


def floats(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])
