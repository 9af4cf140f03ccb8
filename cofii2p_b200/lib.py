"""ctypes binding of libcofi_b200.so (the C ABI declared in include/cofi_b200.h).

There is deliberately no fallback: if the shared library is missing or a symbol cannot be resolved the
import fails loudly, and every op raises when handed a non-CUDA tensor (see ops.py).
Build the library with `python -c "import __graft_entry__ as g; g.build()"` or `cofii2p_b200/csrc/build.sh`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcofi_b200.so")

_vp, _i, _l, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> (restype, argtypes); order mirrors include/cofi_b200.h
SIGNATURES = {
    "cofi_version": (_i, []),
    "cofi_last_error": (ctypes.c_char_p, []),
    "cofi_launch_count": (_l, []),
    "cofi_pack_points": (_i, [_vp, _vp, _l, _i, _l, _vp, _vp]),
    "cofi_kpconv_aggregate": (_i, [_vp, _l, _i, _vp, _vp, _vp, _i, _l, _l, _i, _vp, _i, _f, _f, _vp, _vp, _vp]),
    "cofi_kpconv_aggregate_f16": (_i, [_vp, _l, _i, _vp, _vp, _vp, _i, _l, _l, _i, _vp, _i, _f, _f, _vp, _vp, _vp]),
    "cofi_maxpool_rows": (_i, [_vp, _l, _i, _vp, _i, _l, _l, _i, _vp, _l, _vp]),
    "cofi_maxpool_rows_f16": (_i, [_vp, _l, _i, _vp, _i, _l, _l, _i, _vp, _l, _vp]),
    "cofi_gather_rows": (_i, [_vp, _l, _i, _vp, _l, _l, _l, _i, _vp, _l, _vp]),
    "cofi_half_sample_pyramid": (_i, [_vp, _l, _i, _i, ctypes.c_uint64, _vp, _vp, _vp]),
    "cofi_knn_pyramid_workspace": (_l, [_vp, _i, _i]),
    "cofi_knn_pyramid": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cofi_knn_table_workspace": (_l, [_l, _l, _i]),
    "cofi_knn_table": (_i, [_vp, _l, _vp, _l, _i, _i, _i, _vp, _vp, _vp]),
    "cofi_gemm": (_i, [_vp, _l, _vp, _l, _vp, _l, _l, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "cofi_split_tf32": (_i, [_vp, _l, _i, _i, _vp, _vp]),
    "cofi_debug_x3_profile": (_i, [_vp]),
    "cofi_gemm_f16": (_i, [_vp, _l, _vp, _l, _vp, _l, _l, _i, _i, _vp, _vp, _i, _vp]),
    "cofi_gemm_colstats": (_i, [_vp, _l, _vp, _l, _vp, _l, _l, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "cofi_gemm_colstats_acc": (_i, [_vp, _l, _vp, _l, _vp, _l, _l, _i, _i, _vp, _i, _vp, _vp]),
    "cofi_gemm_f16_colstats": (_i, [_vp, _l, _vp, _l, _vp, _l, _l, _i, _i, _vp, _vp, _vp, _vp]),
    "cofi_gemm_ln": (_i, [_vp, _l, _vp, _l, _vp, _l, _l, _i, _i, _vp, _vp, _vp, _f, _i, _vp, _l, _i, _vp]),
    "cofi_conv2d_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "cofi_norm_rows_workspace": (_l, [_i, _i]),
    "cofi_norm_rows_init": (_i, []),
    "cofi_norm_rows": (_i, [_vp, _l, _l, _i, _i, _i, _vp, _vp, _f, _vp, _l, _i, _vp, _l, _vp, _vp, _vp, _vp]),
    "cofi_norm_rows_pre": (_i, [_vp, _l, _l, _i, _i, _i, _vp, _vp, _f, _vp, _l, _i, _vp, _l, _vp, _vp, _vp]),
    "cofi_affine_rows": (_i, [_vp, _l, _l, _i, _vp, _vp, _vp, _l, _i, _vp, _l, _vp]),
    "cofi_layer_norm_rows": (_i, [_vp, _l, _l, _i, _vp, _vp, _f, _i, _vp, _l, _vp, _l, _vp]),
    "cofi_l2norm_rows": (_i, [_vp, _l, _l, _i, _vp, _l, _vp, _l, _vp]),
    "cofi_colnorm_workspace": (_l, [_i, _i]),
    "cofi_colnorm_rows": (_i, [_vp, _l, _l, _i, _i, _vp, _vp, _l, _vp]),
    "cofi_nchw_to_nhwc": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "cofi_nhwc_to_nchw": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "cofi_maxpool2d_3x3s2_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "cofi_upsample2x_cat_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "cofi_posenc_sine": (_i, [_vp, _l, _i, _i, _vp, _vp, _vp]),
    "cofi_attention": (_i, [_vp, _vp, _vp, _l, _l, _i, _i, _i, _f, _vp, _i, _vp]),
    "cofi_attention_vt": (_i, [_vp, _vp, _vp, _l, _l, _i, _i, _i, _f, _vp, _vp]),
    "cofi_attention_vt_lse": (_i, [_vp, _vp, _vp, _l, _l, _i, _i, _i, _f, _vp, _vp, _vp]),
    "cofi_sim_argmin": (_i, [_vp, _l, _vp, _l, _l, _l, _i, _i, _vp, _vp, _i, _vp]),
    "cofi_sim_argmin_f16": (_i, [_vp, _l, _vp, _l, _l, _l, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "cofi_cast_f16": (_i, [_vp, _l, _l, _i, _vp, _l, _vp]),
    "cofi_cast_f16_bound": (_i, [_vp, _l, _l, _i, _vp, _l, _vp, _vp]),
    "cofi_sim_argmin_exact_workspace": (_l, [_l, _l, _i]),
    "cofi_sim_argmin_exact": (_i, [_vp, _l, _vp, _l, _vp, _l, _vp, _l, _l, _l, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cofi_l2norm_rows_f16": (_i, [_vp, _l, _l, _i, _vp, _l, _vp, _l, _vp]),
    "cofi_select_matches": (_i, [_vp, _vp, _l, _i, _i, _i, _i, _i, _vp, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "cofi_nn_argmin": (_i, [_vp, _l, _vp, _l, _vp, _vp]),
    "cofi_extract_patch": (_i, [_vp, _i, _i, _i, _i, _vp, _l, _vp, _vp, _vp]),
    "cofi_nn_argmin_batched": (_i, [_vp, _l, _vp, _l, _i, _vp, _vp]),
    "cofi_extract_patch_batched": (_i, [_vp, _i, _i, _i, _i, _vp, _l, _vp, _vp, _vp]),
    "cofi_fine_match": (_i, [_vp, _vp, _l, _i, _vp, _vp]),
    # ---- training (backward.cu)
    "cofi_act_bwd": (_i, [_vp, _vp, _l, _i, _vp, _vp]),
    "cofi_rowscale": (_i, [_vp, _l, _i, _vp, _vp, _vp]),
    "cofi_colsum_workspace": (_l, [_i]),
    "cofi_colsum": (_i, [_vp, _l, _l, _i, _vp, _i, _vp, _vp]),
    "cofi_norm_rows_stats": (_i, [_vp, _l, _l, _i, _i, _i, _f, _vp, _vp, _vp]),
    "cofi_norm_rows_bwd_workspace": (_l, [_i, _i]),
    "cofi_norm_rows_bwd": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "cofi_layer_norm_bwd": (_i, [_vp, _vp, _l, _i, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp]),
    "cofi_l2norm_bwd": (_i, [_vp, _vp, _l, _i, _vp, _vp]),
    "cofi_colnorm_bwd_workspace": (_l, [_i, _i]),
    "cofi_colnorm_bwd": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp]),
    "cofi_scatter_add_rows": (_i, [_vp, _l, _i, _vp, _l, _l, _l, _i, _vp, _vp]),
    "cofi_maxpool_rows_bwd": (_i, [_vp, _i, _vp, _i, _l, _l, _i, _vp, _vp, _vp]),
    "cofi_kpconv_aggregate_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _i, _l, _l, _i, _vp, _i, _f, _f, _vp, _vp]),
    "cofi_upsample2x_cat_bwd": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "cofi_maxpool2d_3x3s2_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "cofi_dilate2_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "cofi_extract_patch_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _l, _vp, _vp]),
    "cofi_extract_patch_batched_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _l, _vp, _vp]),
    "cofi_gemm_tn_workspace": (_l, [_l, _i, _i]),
    "cofi_gemm_tn": (_i, [_vp, _l, _vp, _l, _vp, _l, _i, _i, _i, _i, _vp, _vp]),
    "cofi_conv2d_wgrad_workspace": (_l, [_i, _i, _i, _i, _i, _i, _i]),
    "cofi_conv2d_wgrad_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "cofi_attention_fwd_lse": (_i, [_vp, _vp, _vp, _l, _l, _i, _i, _i, _f, _vp, _vp, _vp]),
    "cofi_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _l, _l, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "cofi_attention_bwd_tc_workspace": (_l, [_l, _l, _i, _i, _i]),
    "cofi_attention_bwd_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _l, _l, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cofi_pnp_ransac": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _f, ctypes.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cofi_desc_loss": (_i, [_vp, _vp, _l, _vp, _vp, _l, _vp, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "cofi_overlap_loss": (_i, [_vp, _vp, _l, _i, _i, _i, _vp, _vp, _vp]),
    "cofi_fine_circle_loss": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "cofi_scatter_scaled_rows": (_i, [_vp, _vp, _l, _l, _i, _i, _vp, _f, _vp, _vp]),
    "cofi_adam_step": (_i, [_vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _i, _f, _vp]),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle. Raises ImportError when the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is mandatory (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().cofi_last_error().decode("utf-8", "replace")


def launch_count() -> int:
    return int(load().cofi_launch_count())
