"""ctypes binding of the C-ABI library (include/b200roi.h).

There is deliberately no fallback: if `_C/libb200roi.so` is missing the import raises, and every
entry point raises `B200Error` on a non-zero status.  Build with
`python -m fewshotobjectdetection_imporove_via_text_feature_b200._build` (nvcc, sm_100a).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libb200roi.so")

F32, BF16 = 0, 1
NCHW, NHWC = 0, 1

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t



class Gemm2Desc(ctypes.Structure):
    """struct b200_gemm2_desc (include/b200roi.h), field for field."""
    _fields_ = [("A", c_void_p), ("lda", c_int), ("A2", c_void_p), ("lda2", c_int), ("K2", c_int),
                ("B", c_void_p), ("ldb", c_int), ("M", c_int), ("N", c_int), ("K", c_int),
                ("a_mn", c_int), ("b_mn", c_int), ("conv_c", c_int), ("bias", c_void_p),
                ("residual", c_void_p), ("ld_res", c_int), ("relu", c_int),
                ("mask_act", c_void_p), ("ld_mask", c_int), ("mask_bits", c_void_p), ("ld_mask_bits", c_int),
                ("out_bf16", c_void_p), ("ld_out", c_int), ("out2_bf16", c_void_p), ("ld_out2", c_int),
                ("out_f32", c_void_p), ("ld_out_f32", c_int), ("accumulate", c_int),
                ("bits_out", c_void_p), ("ld_bits_out", c_int), ("rowmean_out", c_void_p), ("ld_rowmean", c_int),
                ("rowsumsq_out", c_void_p), ("ld_rowsumsq", c_int),
                ("row_scale_sumsq", c_void_p), ("ld_row_scale_sumsq", c_int), ("row_scale_parts", c_int), ("row_scale_eps", c_float),
                ("softmax", c_int), ("gate", c_int),
                ("tile_n", c_int), ("max_clusters", c_int), ("epilogue_variant", c_int), ("no_pdl", c_int),
                ("split_k", c_int), ("splitk_workspace", c_void_p), ("splitk_workspace_bytes", c_size_t)]


# name -> (restype, argtypes); mirrors include/b200roi.h declaration by declaration
SIGNATURES = {
    "b200_gemm2": (c_int, [ctypes.POINTER(Gemm2Desc), c_void_p]),
    "b200_gemm2_splitk_workspace_bytes": (c_size_t, [c_int] * 4),
    "b200_abi_version": (c_int, []),
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_set_option": (c_int, [ctypes.c_char_p, c_int]),
    "b200_gdl_affine_fwd": (c_int, [c_void_p] * 4 + [c_int] * 8 + [c_void_p]),
    "b200_gdl_affine_bwd_workspace_bytes": (c_size_t, [c_int] * 4),
    "b200_gdl_affine_bwd": (c_int, [c_void_p] * 3 + [c_float] + [c_void_p] * 3 + [c_int] * 8 + [c_void_p, c_size_t, c_void_p]),
    "b200_roi_align_fwd_workspace_bytes": (c_size_t, [c_int] * 7),
    "b200_roi_align_fwd": (c_int, [c_void_p] * 4 + [c_int] * 8 + [c_float] + [c_int] * 5 + [c_void_p, c_size_t, c_void_p]),
    "b200_roi_align_bwd_workspace_bytes": (c_size_t, [c_int] * 11),
    "b200_roi_align_bwd": (c_int, [c_void_p] * 4 + [c_int] * 8 + [c_float] + [c_int] * 5 + [c_void_p, c_size_t, c_void_p]),
    "b200_roi_align_bwd_plan_bytes": (c_size_t, [c_int] * 8),
    "b200_roi_align_bwd_plan": (c_int, [c_void_p] * 2 + [c_int] * 8 + [c_float] + [c_int] * 2 + [c_void_p, c_size_t, c_void_p]),
    "b200_roi_align_bwd_planned": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p] + [c_int] * 8 + [c_void_p]),
    "b200_softmax_decode_compact_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200_softmax_decode_compact": (c_int, [c_void_p, c_int] + [c_void_p] * 4 + [c_int] * 4 + [c_float] * 5 + [c_void_p] * 6 + [c_int, c_void_p, c_size_t, c_void_p]),
    "b200_batched_nms_workspace_bytes": (c_size_t, [c_int] * 3),
    "b200_batched_nms": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_float, c_int, c_int] + [c_void_p] * 2 + [c_void_p, c_size_t, c_void_p]),
    "b200_gather_detections": (c_int, [c_void_p] * 7 + [c_int] * 2 + [c_void_p] * 4 + [c_void_p]),
    "b200_pcb_cosine_blend": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_float] * 3 + [c_void_p]),
    "b200_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int] + [c_int] * 4 + [c_void_p]),
    "b200_gemm_bf16_ex": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int] + [c_int] * 5 + [c_void_p, c_int, c_void_p]),
    "b200_transpose_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200_colsum_workspace_bytes": (c_size_t, [c_int]),
    "b200_colsum": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "b200_dropout_fwd": (c_int, [c_void_p, c_void_p, c_size_t, c_float, ctypes.c_ulonglong, c_void_p, c_void_p]),
    "b200_residual_layernorm_dropout": (c_int, [c_void_p] * 4 + [c_float, c_int, c_float, ctypes.c_ulonglong, c_void_p, c_void_p,
                                                                  c_int, c_int, c_void_p]),
    "b200_layernorm_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200_layernorm_relu_dropout_bwd": (c_int, [c_void_p] * 5 + [c_float, c_float, ctypes.c_ulonglong] + [c_void_p] * 5 + [c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "b200_layernorm_param_grads": (c_int, [c_void_p] * 5 + [c_float, ctypes.c_ulonglong] + [c_void_p] * 3 + [c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "b200_text_attention_bwd": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 5 + [c_int, c_void_p, c_void_p] + [c_int] * 4 + [c_void_p]),
    "b200_head_losses": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_float] * 5 + [c_void_p, c_void_p, c_void_p]),
    "b200_head_losses_bwd": (c_int, [c_void_p] * 7 + [c_int] * 4 + [c_float] * 5 + [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "b200_kd_loss": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_float] * 2 + [c_void_p, c_void_p]),
    "b200_kd_loss_bwd": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_float] * 2 + [c_void_p, c_int, c_void_p]),
    "b200_skinny_gemm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200_skinny_gemm": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_float, c_void_p,
                                 c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "b200_spatial_mean": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_mean_bwd_relu_mask": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200_add_relu_mask": (c_int, [c_void_p] * 4 + [c_size_t, c_void_p]),
    "b200_spatial_mean_bits": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200_pack_relu_bits": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_mean_bwd_relu_bits": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_add_relu_bits": (c_int, [c_void_p] * 4 + [c_size_t, c_void_p]),
    "b200_sgd_momentum": (c_int, [c_void_p] * 3 + [c_size_t] + [c_float] * 3 + [c_void_p, c_void_p]),
    "b200_text_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int] + [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "b200_residual_layernorm": (c_int, [c_void_p] * 4 + [c_float, c_int] + [c_void_p] * 2 + [c_int] * 2 + [c_void_p]),
    "b200_cast_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200_label_sample_proposals": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_float, c_int, c_int, ctypes.c_ulonglong, c_void_p, c_int, c_int] + [c_void_p] * 7 + [c_void_p]),
    "b200_rpn_select_workspace_bytes": (c_size_t, [c_int] * 4),
    "b200_rpn_select_proposals": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_float] * 2 + [c_void_p] * 4 + [c_void_p, c_size_t, c_void_p]),
    "b200_detector_postprocess": (c_int, [c_void_p] * 7 + [c_int] * 2 + [c_void_p]),
    "b200_gather_rows_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_class_mean_rows_workspace_bytes": (c_size_t, [c_int] * 3),
    "b200_class_mean_rows": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_l2_normalize_rows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p]),
}


class B200Error(RuntimeError):
    pass


# kernels launched per entry point (upper bound for the optional layout-conversion launches is counted where it
# happens); used by bench.py to report `gpu_launches`
KERNELS_PER_CALL = {
    "b200_gdl_affine_fwd": 1, "b200_gdl_affine_bwd": 3, "b200_roi_align_fwd": 1, "b200_roi_align_bwd": 2, "b200_roi_align_bwd_plan": 2, "b200_roi_align_bwd_planned": 1,
    "b200_softmax_decode_compact": 2, "b200_batched_nms": 3, "b200_gather_detections": 1, "b200_pcb_cosine_blend": 1,
    "b200_gemm_bf16": 1, "b200_gemm_bf16_ex": 1, "b200_gemm2": 1, "b200_transpose_bf16": 1, "b200_colsum": 2, "b200_dropout_fwd": 1, "b200_residual_layernorm_dropout": 1,
    "b200_layernorm_relu_dropout_bwd": 4, "b200_layernorm_param_grads": 3, "b200_text_attention_bwd": 1, "b200_head_losses": 1, "b200_head_losses_bwd": 1,
    "b200_sgd_momentum": 1, "b200_spatial_mean": 1, "b200_mean_bwd_relu_mask": 1, "b200_add_relu_mask": 1, "b200_skinny_gemm": 1, "b200_text_attention": 1, "b200_residual_layernorm": 1, "b200_cast_bf16": 1, "b200_l2_normalize_rows": 1, "b200_label_sample_proposals": 1,
    "b200_rpn_select_proposals": 5, "b200_gather_rows_bf16": 1, "b200_kd_loss": 1, "b200_kd_loss_bwd": 1, "b200_spatial_mean_bits": 1, "b200_pack_relu_bits": 1, "b200_mean_bwd_relu_bits": 1, "b200_add_relu_bits": 1, "b200_class_mean_rows": 2, "b200_detector_postprocess": 1,
}
LAUNCHES = 0
# mirror of `g_roi_bwd_impl`'s initial value in csrc/roi_align_bwd.cu (tests restore the option to it;
# tests/test_abi_and_host.py checks the two agree)
ROI_BWD_IMPL_DEFAULT = 2
# bench.py sets PROFILE = {} to have a CUDA event pair recorded around every entry-point call (name -> [(e0, e1, tag)]);
# None (the default) costs nothing
PROFILE = None


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "b200roi: %s is missing — build it with `python -m fewshotobjectdetection_imporove_via_text_feature_b200._build`; "
                "there is no CPU / PyTorch fallback for this path" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def set_option(key, value):
    return call("b200_set_option", key.encode(), int(value))


def call(name, *args, launches=None, tag=None):
    """Invoke an int-returning entry point and raise on failure.  `launches` overrides the per-call kernel count
    when the entry point dispatches to a multi-kernel implementation."""
    global LAUNCHES
    L = lib()
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L, name)(*args)
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1, tag))
    else:
        rc = getattr(L, name)(*args)
    LAUNCHES += KERNELS_PER_CALL.get(name, 0) if launches is None else launches
    if rc != 0:
        raise B200Error("%s failed (%d): %s" % (name, rc, L.b200_last_error().decode("utf-8", "replace")))
    return rc
